/* oracle/gsl_stub -- GAIA_mcmc.c includes this header and uses nothing from it */

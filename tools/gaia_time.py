import time, numpy as np, hb_mcmc_b200 as hb
from hb_mcmc_b200.gaia import GaiaSampler
ctx = hb.Context(0)
D=234.296; d=np.array([7.16094512,-0.0066265000000005,0.0212387299999997,-0.0066558899999999]); e=np.array([0.0230834782584296,0.0367165032930016,0.0586013200752551,0.0086725406204692])
for E in (1, 16, 148*4, 148*16, 148*64):
    s = GaiaSampler(ctx, n_ens=E, seed=5); s.set_data(D,d,e); s.init_random()
    s.run(1000, log=False)
    n = 100000 if E <= 16 else 20000
    t=time.time(); s.run(n, log=False); dt=time.time()-t
    cnt=s.counters()
    print(f"E={E:6d}: {n/dt:10.0f} it/s per ladder, {n*E/dt:12.0f} ladder-it/s, {n*E*20/dt:.3e} walker-steps/s  acc={cnt['accepted'].sum()/cnt['proposed'].sum():.3f} swap={cnt['swaps_accepted'].sum()/cnt['swaps_proposed'].sum():.3f}", flush=True)
    s.close()

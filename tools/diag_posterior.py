"""diagnostic: GPU vs reference posterior summaries with growing sample sizes (is a >3 MCSE gap bias or MCSE under-estimation?)"""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases
import cmdstan_b200 as B
from cmdstan_b200 import analyze as A
from oracle import oracle as O

def summ(d):
    t = torch.tensor(d); K = t.shape[2]
    mean = t.mean(dim=(0, 1)).numpy(); sd = t.reshape(-1, K).std(dim=0).numpy()
    essb = np.array([A.ess_bulk(t[:, :, k]) for k in range(K)])
    essm = np.array([A.ess(A.split_chains(t[:, :, k])) for k in range(K)])
    dev2 = (t - t.mean(dim=(0, 1), keepdim=True)) ** 2
    esss = np.array([A.ess(A.split_chains(dev2[:, :, k])) for k in range(K)])
    evar = dev2.mean(dim=(0, 1)).numpy(); e4 = (dev2 ** 2).mean(dim=(0, 1)).numpy()
    mcse_sd = np.sqrt((e4 - evar ** 2) / esss / evar / 4)
    return mean, sd, sd / np.sqrt(essb), sd / np.sqrt(essm), mcse_sd, essb, essm

def gpu(name, data, C, W, S, seed):
    e = B.Engine(name, data, num_chains=C, random_seed=seed, num_warmup=W, num_samples=S); e.init_chains(); e.run(W + S)
    d = e.fetch_draws(); e.close(); return d

for name, data, cfgs in (("logistic", cases.logistic_data(2000, 8), [(64, 500, 1000), (512, 1000, 2000)]),
                         ("eight_schools_nc", cases.EIGHT_SCHOOLS, [(256, 1000, 1000), (2048, 1000, 2000)])):
    res, _, _ = O.ref_sample(name, data, num_chains=16, seed=777, num_warmup=1000, num_samples=20000, sig_figs=12, num_threads=16)
    ref = np.stack([r[1] for r in res]); cols = [0] + list(range(7, ref.shape[2]))
    mr, sr, mbr, mmr, msr, eb, em = summ(ref[:, :, cols])
    print(name, "ref ess bulk/mean", eb.round(0)[:6], em.round(0)[:6])
    for (C, W, S) in cfgs:
        for seed in (1, 2):
            g = gpu(name, data, C, W, S, seed)
            mg, sg, mbg, mmg, msg, ebg, emg = summ(g[:, :, cols])
            zb = np.abs(mg - mr) / np.sqrt(mbg ** 2 + mbr ** 2); zm = np.abs(mg - mr) / np.sqrt(mmg ** 2 + mmr ** 2); zs = np.abs(sg - sr) / np.sqrt(msg ** 2 + msr ** 2)
            print(f"  C={C} W={W} S={S} seed={seed}: max z_mean(bulk-ess)={zb.max():.2f} max z_mean(mean-ess)={zm.max():.2f} max z_sd={zs.max():.2f}  div={g[:,:,5].sum():.0f} ess_ratio bulk/mean={np.median(ebg/emg):.2f}")
            print("     z_mean:", zm.round(2)); print("     z_sd  :", zs.round(2))

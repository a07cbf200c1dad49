"""Per-phase SM-cycle breakdown of k_resident_step (needs a libsbx built with
-DSBX_PROFILE_PHASES: `NVCC_EXTRA=-DSBX_PROFILE_PHASES python sbsim_b200/build.py --force`)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sbsim_b200 import workloads
B = 32768
env, wl = workloads.make_randomized_env(B, episode_steps=64, histogram=True,
                                        numpy_zone_means=bool(int(os.environ.get("PW", "0"))))   # PW=1: SBX_OPT_NUMPY_MEANS
dev = torch.device("cuda:0")
D = env.observation_spec().shape[0]
obs = torch.zeros(B, D, device=dev); rew = torch.zeros(B, device=dev)
st = torch.zeros(B, dtype=torch.int32, device=dev); dis = torch.zeros(B, device=dev)
env.reset_device(obs, rew, st, dis)
act = torch.rand(40, B, 2, device=dev) * 2 - 1
for i in range(5):
  env.step_device(act[i], obs, rew, st, dis)
torch.cuda.synchronize()
W = 264
env.handle.upload("phase_cycles", np.zeros(W, dtype=np.uint64))
K = 20
for i in range(K):
  env.step_device(act[5 + i], obs, rew, st, dis)
torch.cuda.synchronize()
raw = env.handle.download("phase_cycles", (W,)).astype(np.float64)
cyc = raw[:8]
names = ["prologue (TMA issued)", "wait TMA load", "sweep 1 (+n3)", "sweeps 2..n", "store issue (+gather)",
         "zone sums (warp 0) / PW: tables staged", "barrier after sums / PW: + leaf sums", "PW: levels + means"]
tot = cyc[:8].sum()
for n, c in zip(names, cyc):
  print(f"{n:28s} {c / (B * K):9.0f} cycles/CTA  {100 * c / tot:5.1f}%")
print(f"total {tot / (B * K):.0f} cycles/CTA = {tot / (B * K) / 1.965e3:.2f} us at 1.965 GHz")

# the probe CTA (building B/2, last step): cycles from CTA start to each warp ARRIVING at the
# barrier that ends sweep k, and to LEAVING it (= the slowest warp's arrival)
probe = raw[8:8 + 16 * 8].reshape(16, 8)
for k in range(3):
  arr, leave = probe[:, 2 * k], probe[:, 2 * k + 1]
  if leave.max() == 0:
    continue
  print(f"sweep {k + 1}: arrivals " + " ".join(f"{int(a):6d}" for a in arr) + f" | released at {int(leave.max())}")

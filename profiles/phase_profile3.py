"""Per-phase SM-cycle breakdown of k_resident_step3 (needs a libsbx built with
-DSBX_PROFILE_PHASES, e.g. NVCC_EXTRA=-DSBX_PROFILE_PHASES SBX_LIB_OUT=... python sbsim_b200/build.py --force;
run with SBX_LIB pointing at it)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sbsim_b200 import workloads
B = 32768
env, wl = workloads.make_randomized_env(B, episode_steps=64, histogram=True)
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
i0 = env.handle.info()
for i in range(K):
  env.step_device(act[5 + i], obs, rew, st, dis)
torch.cuda.synchronize()
i1 = env.handle.info()
raw = env.handle.download("phase_cycles", (W,)).astype(np.float64)
cyc = raw[:8]
names = ["load wait (TMA + lists)", "pattern table, tile load", "n3 + first exchange", "sweep 1", "sweeps 2..n", "write-back + zone sums", "outputs + store drain", "-"]
tot = cyc.sum()
print(f"mean sweeps/step {(i1.sweeps_total - i0.sweeps_total) / (B * K):.2f}")
for n, c in zip(names, cyc):
  print(f"{n:28s} {c / (B * K):9.0f} cycles/building  {100 * c / tot:5.1f}%")
print(f"total {tot / (B * K):.0f} cycles/building = {tot / (B * K) / 1.965e3:.2f} us at 1.965 GHz")
probe = raw[8:8 + 16 * 8].reshape(16, 8)
for k in range(3):
  arr, leave = probe[:, 2 * k], probe[:, 2 * k + 1]
  if leave.max() == 0:
    continue
  print(f"sweep {k + 1}: arrivals " + " ".join(f"{int(a):6d}" for a in arr) + f" | released at {int(leave.max())}")

"""Short program for ncu: 2 persistent launches (T time steps each) then 2 split steps, C2 shapes."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vjf_b200 import _lib
from vjf_b200.model import VJF

B = int(os.environ.get("PB", 4096)); T = int(os.environ.get("PT", 4))
D, d, R, H = int(os.environ.get("PD", 200)), int(os.environ.get("Pd", 3)), int(os.environ.get("PR", 50)), [int(os.environ.get("PH", 64))]
torch.manual_seed(0)
m = VJF.make_model(D, d, 0, R, H, os.environ.get("PLIK", "poisson"), max_trials=B)
y = torch.poisson(torch.full((T, B, D), 0.5, device="cuda"))
for _ in range(2):
    mu, lv, ls = m.run(y)
torch.cuda.synchronize()
lib = _lib.load()
p = lambda t: C.c_void_p(0 if t is None else t.data_ptr())
om, ol, loss = torch.empty(B, d, device="cuda"), torch.empty(B, d, device="cuda"), torch.empty(4, device="cuda")
fl = _lib.FLAG_SGD | _lib.FLAG_UPDATE
for t in range(2):
    _lib.check(lib.vjf_step_phase_a(m._h, B, B, p(y[t]), 0, None, p(mu[-1]), p(lv[-1]), None, 0, t, 0, fl, p(om), p(ol), None))
    _lib.check(lib.vjf_step_phase_b(m._h, B, fl, m.lr, p(loss), None))
torch.cuda.synchronize()
print("ok", ls[-1].tolist(), loss.tolist(), m.status())

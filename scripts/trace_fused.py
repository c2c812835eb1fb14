import os, sys, ctypes
os.environ.setdefault("KVC_FUSED_COOP", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import _native
dev = torch.device("cuda:0"); L = 32768
Ks = [torch.randn(1, 8, L, 128, device=dev).to(torch.bfloat16) for _ in range(4)]
Q = torch.randn(1, 32, L, 128, device=dev).to(torch.bfloat16)
for i in range(8):
    s = _native.window_scores(Q, Ks[i % 4], 32, "sum", 3)
torch.cuda.synchronize()
raw = ctypes.CDLL(_native.LIB_PATH); buf = (ctypes.c_longlong * 256)(); raw.kvc_debug_ftrace(buf)
t = list(buf); t0 = t[0]; r = lambda i: t[i] - t0
print("setup->0; Q staged", r(1), "| pass1 done", r(2), "| stats published", r(3), "| barrier passed", r(4), "| crow ready", r(5), "| end", r(6))
print("pass1 wg0 tfull:", [r(16 + j) for j in range(0, 15, 3)])
print("pass1 wg0 done :", [r(64 + j) for j in range(0, 15, 3)])
print("pass2 wg0 tfull:", [r(112 + j) for j in range(15, 30, 3)])

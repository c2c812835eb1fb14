import os, sys, ctypes
os.environ["KVC_TC_DEBUG"] = sys.argv[1] if len(sys.argv) > 1 else "4"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import _native
dev = torch.device("cuda:0"); L = 32768
K = torch.randn(1, 8, L, 128, device=dev).to(torch.bfloat16)
Q = torch.randn(1, 32, L, 128, device=dev).to(torch.bfloat16)
for _ in range(3):
    s = _native.window_scores(Q, K, 32, "sum", 2)
torch.cuda.synchronize()
raw = ctypes.CDLL(_native.LIB_PATH)
buf = (ctypes.c_longlong * 512)()
print("rc", raw.kvc_debug_trace(buf))
t = list(buf); t0 = t[0]
rel = lambda i: t[i] - t0
print("pass-2 kernel (last launch), cycles relative to entry of CTA(0,0):")
print(" setup done", rel(1), " Q staged (MMA warp)", rel(2), " teardown", rel(3))
print(" TMA issued  :", [rel(32 + i) for i in range(15)])
print(" kfull (MMA) :", [rel(64 + i) for i in range(15)])
print(" tfull wg0   :", [rel(96 + j) for j in range(0, 15, 2)])
print(" tfull wg1   :", [rel(96 + j) for j in range(1, 15, 2)])
print(" epi done wg0:", [rel(128 + j) for j in range(0, 15, 2)])
print(" epi done wg1:", [rel(128 + j) for j in range(1, 15, 2)])
print(" loads done wg0:", [rel(160 + j) for j in range(0, 15, 2)])
print(" conv done wg0 :", [rel(192 + j) for j in range(0, 15, 2)])
print(" loads done wg1:", [rel(160 + j) for j in range(1, 15, 2)])
print(" conv done wg1 :", [rel(192 + j) for j in range(1, 15, 2)])
print(" token got wg0 :", [rel(224 + j) for j in range(0, 15, 2)])
print(" token got wg1 :", [rel(224 + j) for j in range(1, 15, 2)])

import os, sys, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from kvcache_factory_b200 import SnapKVCluster
dev = torch.device("cuda:0")
L = 4096
K = torch.randn(1, 8, L, 128, device=dev).to(torch.bfloat16); V = torch.randn(1, 8, L, 128, device=dev).to(torch.bfloat16)
Q = torch.randn(1, L, 32, 128, device=dev).to(torch.bfloat16).transpose(1, 2)
cl = SnapKVCluster(32, 128, 7, "avgpool")
for i in range(20): cl.update_kv(K, Q, V, None, 4)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for i in range(2000): cl.update_kv(K, Q, V, None, 4)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)

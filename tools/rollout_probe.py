"""Forward-only stack vs training forward of the Water-3D batch, un-captured, for an ncu launch list:
ncu --metrics gpu__time_duration.sum ... python tools/rollout_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
data, hp = bench.make_workload("water3d", 0, 0)
sb = bench.StepBench(data, hp, dev, use_cuda_graph=False)
t = sb.dev_in
kw = dict(node_feat=t["node_feat"], node_loc=t["loc_0"], node_vel=t["vel_0"], edge_index=t["edge_index"],
          data_batch=t["batch"], loc_mean=t["loc_mean"], edge_attr=t["edge_attr"])
m = sb.model
m.eval()
for keep in (False, True, False, True):
    m.eval_keeps_graph = keep
    with torch.no_grad() if not keep else torch.enable_grad():
        for _ in range(2):
            m(**kw)
    torch.cuda.synchronize()
    torch.zeros(1, device=dev).add_(1)          # marker kernel between the sections
    torch.cuda.synchronize()

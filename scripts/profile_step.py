"""One profiled training step of the bench workload, bracketed by cudaProfilerStart/Stop (use with
`ncu --profile-from-start off ...`).  Usage: python scripts/profile_step.py [config] [batch] [warm steps] [nvit|orig|kohonen]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nvit_b200 import ViT, ViTConfig, Trainer
import bench

name = sys.argv[1] if len(sys.argv) > 1 else "b16"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
warm = int(sys.argv[3]) if len(sys.argv) > 3 else 2
variant = sys.argv[4] if len(sys.argv) > 4 else "nvit"
cfg = ViTConfig(**bench.config_dict(name, variant))
torch.manual_seed(0)
model = ViT(cfg).cuda().train()
tr = Trainer(model)
g = torch.Generator().manual_seed(1234)
X = torch.randn(B, 3, cfg.image_size, cfg.image_size, generator=g).cuda()
y = torch.randint(0, cfg.num_classes, (B,), generator=g).cuda()
for _ in range(warm):
    tr.step(X, y)
torch.cuda.synchronize()
torch.cuda.profiler.start()
tr.step(X, y)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled one step; loss", float(tr.loss_buf), "launches", tr.total_launches)

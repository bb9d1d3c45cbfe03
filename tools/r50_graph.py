"""ResNet3D-50 (and -18) training step in CUDA-graph mode with whatever multimodal_ad_b200 package is first on sys.path
(bisecting tool: run from an exported older tree with PYTHONPATH=<tree>).  usage: r50_graph.py [depth] [batch]"""
import json, os, sys
import torch, torch.nn as nn
from multimodal_ad_b200.models.Resnet3D import generate_model

depth = int(sys.argv[1]) if len(sys.argv) > 1 else 50
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = generate_model(model_depth=depth, input_W=128, input_H=128, input_D=128, nb_class=2, pretrain_path=None, dropout_rate=0.5, device=dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-5, weight_decay=1e-4, fused=True, capturable=True)
crit = nn.CrossEntropyLoss()
x = torch.rand(batch, 1, 128, 128, 128, device=dev); y = torch.randint(0, 2, (batch,), device=dev)
def step():
    loss = crit(model(x), y); opt.zero_grad(set_to_none=True); loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0); opt.step(); return loss
for _ in range(3): step()
torch.cuda.synchronize()
opt.zero_grad(set_to_none=True)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    l = crit(model(x), y); l.backward(); torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0); opt.step()
for _ in range(3): g.replay()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): g.replay()
b.record(); b.synchronize()
import multimodal_ad_b200
print(json.dumps(dict(tree=os.path.dirname(multimodal_ad_b200.__file__), depth=depth, batch=batch, ms=round(a.elapsed_time(b) / 10, 3))))

"""Host-side issue time vs device time of the headline train step (is the step launch bound?)."""
import os, sys, time, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import quadtree_oracle as O
from qtcnn_b200 import models as M
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = M.QuadtreeCNN(num_classes=8).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-4, fused=True)
x, nf, y = (t.to(dev) for t in O.synthetic_batch(256, 1234))
def step():
    opt.zero_grad(set_to_none=True)
    loss = F.cross_entropy(model(x, nf), y)
    loss.backward()
    opt.step()
for _ in range(5): step()
torch.cuda.synchronize()
for n in (1, 20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(n): step()
    t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
    print(f"steps {n}: host issue {1e3*(t1-t0)/n:.2f} ms/step, device {e0.elapsed_time(e1)/n:.2f} ms/step")
# host-only cost of forward / backward / optimizer
t0 = time.perf_counter(); opt.zero_grad(set_to_none=True); loss = F.cross_entropy(model(x, nf), y); t1 = time.perf_counter()
loss.backward(); t2 = time.perf_counter(); opt.step(); t3 = time.perf_counter(); torch.cuda.synchronize()
print(f"host: forward {1e3*(t1-t0):.2f} ms, backward {1e3*(t2-t1):.2f} ms, optimizer {1e3*(t3-t2):.2f} ms (queue empty at start)")

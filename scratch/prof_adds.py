"""Which torch elementwise ops run in a c3 training step, and from which Python lines (torch.profiler with stacks)."""
import sys; sys.path.insert(0, ".")
import torch, collections
import xequinet_b200 as xb
from oracle import xpainn_oracle as orc
import bench
from torch.profiler import profile, ProfilerActivity
dev = "cuda"; cfg = orc.CONFIG_DEFAULT
model = xb.resolve_model("xpainn", **cfg.model_kwargs()); model.load_state_dict(orc.synthetic_state_dict(cfg, 1234), strict=False)
model = model.to(dev).train(); params = list(model.parameters()); opt = torch.optim.AdamW(params, lr=5e-4, fused=True)
tr = xb.NeighborTransform(cfg.cutoff)
d0 = {k: v.to(dev) for k, v in bench.make_batch("c3", 256, 0).items()}
def step():
    d = tr({k: d0[k] for k in ["pos", "atomic_numbers", "batch", "ptr", "target_energy", "target_forces"]})
    out = model(d, compute_forces=True); loss = bench.loss_fn(out, d, True)
    opt.zero_grad(set_to_none=True); loss.backward(); opt.step()
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
    step(); torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.name in ("aten::add", "aten::add_", "aten::copy_", "aten::mul", "aten::clone", "aten::contiguous", "aten::cat", "aten::zeros", "aten::fill_", "aten::zero_", "aten::sum", "aten::neg", "aten::zeros_like") and e.device_time_total > 0:
        st = [s for s in (e.stack or []) if "xequinet_b200" in s or "bench.py" in s]
        seqs = " <- ".join(s.split("/")[-1][:60] for s in st[:2]) or ("autograd:" + ((e.stack or ["?"])[0][-60:]))
        key = (e.name, str(e.input_shapes)[:60], seqs)
        agg[key][0] += 1; agg[key][1] += e.device_time_total
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(v[0], round(v[1], 1), k)

"""Per-kernel device times of the training step under torch.profiler (CUPTI activity tracing: warm caches, no replay, no
serialisation) -- the complement of the ncu launch list, whose per-launch times are cold-cache."""
import importlib, os, sys, json, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
pkg = importlib.import_module("visual-question-answering_b200")
dev = torch.device("cuda", 0)
use_graph = "--graph" in sys.argv
st = bench.Stepper(pkg, dev, 160, 1, None, use_graph)
st.warm(3)
if use_graph:
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        st.warm(3)
    torch.cuda.current_stream().wait_stream(s)
    st.capture()
for i in range(3):
    st.step(i)
torch.cuda.synchronize()
N = 6
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    for i in range(N):
        st.step(i)
    torch.cuda.synchronize()
agg = collections.OrderedDict()
order = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1; a[1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
tot = sum(v[1] for v in agg.values())
print(f"# torch.profiler, {N} steps ({'graph replay' if use_graph else 'eager'}): {tot / N:.1f} us of kernel time per step")
print("| us/step | share | launches/step | kernel |\n|---:|---:|---:|---|")
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"| {us / N:.1f} | {100 * us / tot:.1f}% | {n / N:.1f} | `{k[:120]}` |")

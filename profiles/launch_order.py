"""Every kernel launch of one training step, in stream order, with its exclusive device time (CUPTI through torch.profiler; eager, PDL off)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
sys.argv = [sys.argv[0]]
import bench
pkg = importlib.import_module("visual-question-answering_b200")
importlib.import_module("visual-question-answering_b200.dp")
dev = torch.device("cuda:0")
st = bench.Stepper(pkg, dev, 160, 1, None, False)
st.warm(3)
pkg._lib.set_option("pdl", "0")
st._step_body(st.slots[0]); torch.cuda.synchronize()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    st._step_body(st.slots[1])
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
tot = 0.0
for i, e in enumerate(evs):
    name = e.name.replace("hca::(anonymous namespace)::", "").replace("void ", "")
    tot += e.device_time
    print(f"{i:3d} {e.device_time:8.1f} us  {name[:110]}")
print("total", tot)

"""Phase timing of ssg_post_processing_batched (64 images, 8 confident instances each) with CUDA events."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crog_b200 import synth
from crog_b200 import _lib as L
from crog_b200.utils import grasp_eval as GE

dev = torch.device("cuda", 0)
cfg = synth.ssg_cfg()
B = 64
ods = [synth.make_ssg_output_dict(cfg, n_confident=8, seed=100 + i) for i in range(B)]
od = {k: torch.cat([o[k] for o in ods]).to(dev) for k in ("protos", "cls_pred", "box_pred", "ins_coef_pred", "grasp_coef_pred")}
od["anchors"] = ods[0]["anchors"]
for _ in range(3):
    GE.ssg_post_processing_batched(cfg, od, (480, 640))
torch.cuda.synchronize()
ts = []
for _ in range(5):
    t0 = time.perf_counter(); GE.ssg_post_processing_batched(cfg, od, (480, 640)); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print("whole call (wall, incl. sync): median %.2f ms" % sorted(ts)[2])
# phases: monkey-patch the library calls with event brackets
lib = L.lib()
marks = []
def wrap(name):
    fn = getattr(lib, name)
    def w(*a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s = torch.cuda.ExternalStream(a[-1]) if a[-1] else torch.cuda.current_stream()
        e0.record(s); r = fn(*a); e1.record(s); marks.append((name, e0, e1)); return r
    return w
class Lib:
    def __getattr__(self, n):
        return wrap(n) if n.startswith("crog_") else getattr(lib, n)
GE.L.lib = lambda: Lib()
t0 = time.perf_counter(); GE.ssg_post_processing_batched(cfg, od, (480, 640)); torch.cuda.synchronize(); print("instrumented wall %.2f ms" % ((time.perf_counter() - t0) * 1e3))
import collections
agg = collections.defaultdict(lambda: [0, 0.0])
for n, a, b in marks:
    agg[n][0] += 1; agg[n][1] += a.elapsed_time(b)
for n, v in agg.items(): print("%-28s x%3d  sum %.3f ms" % (n, v[0], v[1]))
first = marks[0][1]; last = marks[-1][2]
print("first launch -> last kernel end: %.3f ms" % first.elapsed_time(last))

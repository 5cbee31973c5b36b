#!/usr/bin/env python
"""Per-phase timing of the persistent decode-step kernel (WHISPER_B200_STEP_TRACE=1): for every grid barrier, when the first /
last CTA arrived and when CTA 0 was released.  Diagnostic tool."""
import os, sys
os.environ["WHISPER_B200_STEP_TRACE"] = sys.argv[3] if len(sys.argv) > 3 else "1"      # row groups of the launches to trace
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "godot-whisper_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np
import bench, whisper_b200 as wb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
blob, _ = bench.model_bytes_for(sys.argv[2] if len(sys.argv) > 2 else "tiny.en")
lib = wb.load_library(); ctx = wb.Context(blob, device=0)
chunks = bench.load_inputs(B)
p = wb.host_params(lib, max_tokens=0, entropy_thold=2.4, temperature_inc=0.0, n_threads=4)
for _ in range(2):
    assert ctx.full_batch(p, chunks) == 0
tr = ctx.read_stage(8, np.uint64).reshape(-1, 112, 8).astype(np.int64)
G = tr.shape[0]
if len(sys.argv) > 3 and int(sys.argv[3]) > 1:
    G = G // int(sys.argv[3]); tr = tr[:G]        # first row group only
n_ph = int((tr[0, :, 0] > 0).sum())
t0 = tr[:, 0, 0].min()
names = [f"L{l}.{k}" for l in range(n_ph // 8) for k in ("qkv", "self", "wo", "cq", "cross", "wco", "fc1", "fc2")] + ["logits"]
prev_rel = t0
tr[:, n_ph - 1, 1] = tr[:, n_ph - 1, 0].max()      # the last phase ends the kernel: no barrier after it
print(f"grid {G}, {n_ph} phases, step total {(tr[:, n_ph - 1, 0].max() - t0) / 1e3:.1f} us  (last step of a {B}-chunk batch)")
for ph in range(n_ph):
    arr = tr[:, ph, 0]; rel = tr[:, ph, 1]
    print(f"{names[ph] if ph < len(names) else ph:10s} first arrive {(arr.min() - prev_rel) / 1e3:7.2f}  median {(np.median(arr) - prev_rel) / 1e3:7.2f}  "
          f"last arrive {(arr.max() - prev_rel) / 1e3:7.2f} (cta {int(arr.argmax()):3d})  released +{(rel.min() - arr.max()) / 1e3:5.2f}..+{(rel.max() - arr.max()) / 1e3:5.2f} us")
    c = int(arr.argmax())
    if tr[c, ph, 2] > 0:
        st = tr[c, ph]; base = tr[c, ph - 1, 1] if ph > 0 else t0
        print(f"           slowest cta {c}: released->staged {(st[2] - base) / 1e3:5.2f}  ->data landed {(st[3] - st[2]) / 1e3:5.2f}  "
              f"->first job done {(st[4] - st[3]) / 1e3:5.2f}  ->arrive {(st[0] - st[4]) / 1e3:5.2f}   cycles waiting {st[5]} computing {st[6]} issuing {st[7]}")
    prev_rel = rel.min()
ctx.close()

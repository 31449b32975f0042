"""One device-resident search of a whole BASELINE config (default C5: ONE launch of the segment
kernel) -- the launch `bench.py`'s device leg times; used for the ncu capture behind profiles/traffic.json.
    python tests/gpu_c5_device.py [C5|C2|C4] [reps]"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import __graft_entry__ as g  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "C5"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
corpus = g.load_submodule("corpus")
data = np.frombuffer(corpus.generate_cached(cfg) if cfg == "C5" else corpus.generate(cfg), dtype=np.uint8)
pkg = g.load_package()
n, W, t = len(data), 8192, 15
dev = torch.device("cuda", 0)
d_x = torch.zeros(pkg.required_bytes(n, W), dtype=torch.uint8, device=dev)
d_x[:n].copy_(torch.from_numpy(data))
d_l = torch.empty(n, dtype=torch.uint8, device=dev)
s = torch.cuda.current_stream()
for rep in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(s)
    pkg.search_device(0, d_x.data_ptr(), n, W, t, d_l.data_ptr(), None, s.cuda_stream)
    e1.record(s)
    torch.cuda.synchronize()
    print(f"{cfg} device-resident search rep {rep}: {e0.elapsed_time(e1):.3f} ms", flush=True)
sha = hashlib.sha256(d_l.cpu().numpy().tobytes()).hexdigest()
want = json.loads((ROOT / "tests" / "golden" / "tables.json").read_text()).get(cfg, {}).get("lstar_sha256")
print("table sha256", sha, "== recorded" if sha == want else f"!= recorded {want}")

"""Phase timeline of the tcgen05 GWB synthesis kernel (ptar_debug_i8_timestamps): clock64() stamps of the CTAs of
r-block 0, printed as per-phase medians in SM clocks.   python tools/i8_timeline.py   (GPU box)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from pta_replicator_b200 import _cabi, synthetic  # noqa: E402
from pta_replicator_b200.engine import PulsarBatch  # noqa: E402

psrs, noise = synthetic.make_ng15_like("full")
b = PulsarBatch(psrs)
synthetic.ng15_recipe(b, noise, white=False, ecorr=False, red=False)
st = b.compile()
R = 1000
b.generate(R, seed=1)
buf = torch.zeros(int(st["i8_tiles"].shape[0]) * 8, dtype=torch.int64, device=b.device)
_cabi.check(_cabi.lib().ptar_debug_i8_timestamps(buf.data_ptr()))
b.generate(R, seed=1)
torch.cuda.synchronize()
_cabi.check(_cabi.lib().ptar_debug_i8_timestamps(None))
t = buf.view(-1, 8).cpu().numpy().astype(np.float64)
names = ["setup (barriers, TMEM alloc)", "loads issued (producer done)", "first stage landed", "MMAs issued", "accumulators complete",
         "epilogue done"]
base = t[:, 0]
print("tiles", len(t), "k-chunks: median", np.median(t[:, 7]), "mean", t[:, 7].mean())
for i, n in enumerate(names, start=1):
    d = t[:, i] - base
    print(f"  {n:32s} median {np.median(d):9.0f}  p10 {np.quantile(d, 0.1):9.0f}  p90 {np.quantile(d, 0.9):9.0f} clk since CTA start")
per = (t[:, 5] - t[:, 3]) / np.maximum(t[:, 7], 1)
print(f"  mainloop clk per 32-column chunk (first stage landed -> accumulators complete): median {np.median(per):.0f} "
      "(MMA floor 416 for one CTA alone; two CTAs share the SM's tensor core)")
print(f"  epilogue: median {np.median(t[:, 6] - t[:, 5]):.0f} clk")

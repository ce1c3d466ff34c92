"""clock64 timeline of one CTA of the AST attention kernel (key tiles 8..11): where a key-tile period goes.
python scripts/attn_timeline.py"""
import ctypes as C
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from amuse_b200.engine import Engine          # noqa: E402
from oracle import weights as W               # noqa: E402

eng = Engine("cuda:0")
eng.load_state_dict("ast", W.ast_state_dict(depth=1))
eng.finalize()
fb = (torch.randn(32, 1024, 128, generator=torch.Generator().manual_seed(0)) * 0.5).cuda()
eng.ast_features(fb)
torch.cuda.synchronize()
eng._check(eng.lib.amuse_debug_attn_profile(eng._h, 1, None, 0))
eng.ast_features(fb)
buf = (C.c_int64 * 64)()
eng._check(eng.lib.amuse_debug_attn_profile(eng._h, 0, buf, 64))
st = list(buf)
t0 = min(x for x in st if x > 0)
names = {0: "sm0 wait S", 1: "sm0 S done", 2: "sm0 S in regs", 3: "sm0 P computed", 4: "sm0 P stored+signalled",
         8: "mma wait P0", 9: "mma P0 ready", 10: "mma PV0 issued", 11: "mma S0' issued",
         12: "mma wait P1", 13: "mma P1 ready", 14: "mma PV1 issued", 15: "mma S1' issued"}
ev = []
for j in range(4):
    for s, n in names.items():
        v = st[j * 16 + s]
        if v:
            ev.append((v - t0, j + 8, n))
for t, j, n in sorted(ev):
    print(f"{t:8d}  j={j}  {n}")

"""Developer timing probe: decode of 64 clips (two 32-clip passes) with the tcgen05 and the fp32 attention kernel."""
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from amuse_b200.engine import Engine
from oracle import weights as W
eng = Engine("cuda:0")
eng.load_state_dict("vae", W.motionprior_state_dict()); eng.finalize()
z = torch.randn(64, 128, generator=torch.Generator().manual_seed(0)).cuda()
for B in (64, 32):
    ts = []
    for i in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = eng.decode(z[:B]); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print("decode B=%%d ms: %%s  median %%.3f" %% (B, " ".join("%%.3f" %% t for t in ts), sorted(ts)[4]))
''' % str(ROOT)
for env in ({}, {"AMUSE_ATTN_FFMA": "1"}, {"AMUSE_DECODE_GRAPH": "0"}):
    r = subprocess.run([sys.executable, "-c", CHILD], env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
    print(env, "\n", r.stdout.strip(), r.stderr.strip()[-500:] if r.returncode else "", flush=True)

"""Developer timing probe (not the bench contract): device-time of the sampler pieces + the
in-kernel cycle profile of one denoiser step.  Usage: python scripts/quick_bench.py [B]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from amuse_b200.engine import Engine          # noqa: E402
from oracle import weights as W              # noqa: E402


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    eng = Engine("cuda:0")
    eng.load_state_dict("denoiser", W.denoiser_state_dict())
    eng.load_state_dict("vae", W.motionprior_state_dict())
    eng.finalize()
    g = torch.Generator().manual_seed(0)
    l0, con, emo, sty = (torch.randn(B, d, generator=g).cuda() for d in (128, 256, 256, 256))
    noise = torch.randn(1000, B, 128, generator=g).cuda()
    for n, samp in ((50, "ddim"), (1000, "ddpm")):
        ms = timeit(lambda: eng.denoise(l0, con, emo, sty, n_steps=n, sampler=samp,
                                        step_noise=noise if samp == "ddpm" else None), n=3, warm=1)
        print(f"denoise B={B} {samp}{n}: {ms:.3f} ms  ({ms / n * 1000:.2f} us/step)")
    z = eng.denoise(l0, con, emo, sty, n_steps=50)
    ms = timeit(lambda: eng.decode(z))
    print(f"decode B={B}: {ms:.3f} ms   ({B * 1.7595 / ms:.2f} TFLOP/s algorithmic)")
    ms = timeit(lambda: eng.diffusion_backward(l0, con, emo, sty, n_steps=1000, sampler="ddpm", step_noise=noise), n=3, warm=1)
    print(f"diffusion_backward B={B} ddpm1000: {ms:.3f} ms  -> {B * 300 / ms * 1000:.0f} frames/s")
    ms = timeit(lambda: eng.diffusion_backward(l0, con, emo, sty, n_steps=50, sampler="ddim"), n=3, warm=1)
    print(f"diffusion_backward B={B} ddim50:   {ms:.3f} ms  -> {B * 300 / ms * 1000:.0f} frames/s")
    # in-kernel cycle stamps of step 3
    eng.profile_arm(3)
    eng.denoise(l0, con, emo, sty, n_steps=8)
    st = eng.profile_read(512)
    names = ["skip", "qkv", "attn", "oproj+wait", "sum+ln1", "ffn1", "ffn2+wait", "sum+ln2"]
    print("step cycles total:", st[92] - st[0], " build-x:", st[1] - st[0])
    for l in range(9):
        base = 2 + l * 10
        prev = st[1] if l == 0 else st[2 + (l - 1) * 10 + 7]
        row = []
        for j in range(8):
            row.append(st[base + j] - (prev if j == 0 else st[base + j - 1]))
        print(f"layer {l}: " + " ".join(f"{n}={c}" for n, c in zip(names, row)))
    print("final+update:", st[92] - st[2 + 8 * 10 + 7])
    if len(st) >= 400 and any(st[300:380]):   # issuer: cycles held by the weights | then waiting for the B operand, per stage
        kinds = ["qkv", "wo", "w1", "w2"]
        names = [f"L{i // 4}.{kinds[i % 4]}" for i in range(20)] + \
                [f"L{5 + i // 5}.{(['sk'] + kinds)[i % 5]}" for i in range(20)]
        late = [(names[i], st[300 + 2 * i], st[301 + 2 * i]) for i in range(40)]
        print("issuer per stage (weights-wait | B-wait): " + "  ".join(f"{n}={a}|{b}" for n, a, b in late))
        print("stages held up by their weights (B-wait < 60):", [n for n, a, b in late if b < 60],
              " total weights-wait", sum(a for _, a, _ in late))
    if not st[128]:
        print("layer0 exchange waits: out_proj", st[5] - st[105], " ffn2", st[8] - st[108])
    if len(st) >= 512 and st[128]:   # tensor-core kernel: per-warp stamps inside the out_proj and FFN1 stages of layer 1
        t0 = st[128]
        iss = [st[108] - t0, st[109] - t0, st[112] - t0, st[113] - t0]
        print(f"issuer (relative to warp 0 entering out_proj): bready {iss[0]} issued {iss[1]} | ffn1: bready {iss[2]} issued {iss[3]}")
        names = ["start", "mma-done", "drained", "sent", "recvd", "ln-pre-bar", "ln-post-bar", "ln-done", "b-written",
                 "ffn1-start", "ffn1-d0", "ffn1-gelu0", "ffn1-d1", "ffn1-gelu1"]
        print("warp(r,q) " + " ".join(f"{n:>11}" for n in names))
        for w in range(20):
            row = st[128 + w * 16: 128 + w * 16 + 14]
            print(f"   ({w >> 2},{w & 3})  " + " ".join(f"{v - t0:11d}" for v in row))
    elif st[112]:   # library built with -DAMUSE_FINE_PROF: inside the stages of layer 1 (thread 0)
        print(f"fine qkv : acquire={st[120] - st[12]} gemm={st[121] - st[120]} park+sync={st[122] - st[121]} "
              f"gather={st[123] - st[122]} sync={st[13] - st[123]}")
        print(f"fine oprj: acquire={st[124] - st[14]} gemm={st[125] - st[124]} park+sync={st[126] - st[125]} "
              f"gather+send+wait={st[15] - st[126]} sum+ln+sync={st[16] - st[15]}")
        print(f"fine ffn1: acquire={st[112] - st[16]} gemm={st[113] - st[112]} park+sync={st[114] - st[113]} "
              f"gather+gelu={st[115] - st[114]} sync={st[17] - st[115]}")
        print(f"fine ffn2: acquire={st[116] - st[17]} gemm={st[117] - st[116]} park+sync={st[118] - st[117]} "
              f"gather+send+wait={st[18] - st[118]} sum+ln+sync={st[19] - st[18]}")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Bring-up diagnostics for the GPU box: every step runs in its own process under a timeout, prints numbers instead of
asserting, and the whole log is written to gpurun_out/bringup.log.  TEST TOOLING (uses oracle/ as the checker)."""
import os, sys, subprocess, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "godot-whisper_b200"))
import numpy as np


def step_gemm(engine):
    import whisper_b200 as wb
    rng = np.random.default_rng(0)
    shapes = [(128, 128, 64), (128, 128, 256), (384, 1500, 384), (64, 1500, 1504), (1500, 1500, 64), (384, 300, 240),
              (1536, 1500, 384), (384, 1500, 1536), (51864, 40, 384), (1152, 9, 384)]
    for (M, N, K) in shapes:
        A = (rng.standard_normal((M, K)) * 0.5).astype(np.float16)
        B = (rng.standard_normal((N, K)) * 0.5).astype(np.float16)
        ref = B.astype(np.float32) @ A.astype(np.float32).T
        t0 = time.time()
        out, ms = wb.gemm_f16(A, B, engine=engine, iters=3)
        err = np.abs(out - ref).max()
        print(f"gemm engine={engine} M={M} N={N} K={K}: max_abs_err={err:.3e} ref_max={np.abs(ref).max():.2f} ms/iter={ms:.4f} "
              f"TF={2.0*M*N*K/ms/1e9:.2f} wall={time.time()-t0:.2f}s", flush=True)


def _load_both():
    import whisper_b200 as wb
    from oracle import ref_lib
    model = open(ref_lib.tiny_en_model_path(), "rb").read()
    pcm = ref_lib.read_wav_f32(ref_lib.jfk_wav_path())
    rlib = ref_lib.load()
    ref = ref_lib.RefSession(rlib, model, use_gpu=False)
    log = []
    wb.set_log_sink(wb.load_library(), log)
    ctx = wb.Context(model)
    return wb, ref_lib, ref, ctx, pcm, log


def cmp(name, a, b):
    a = np.asarray(a, dtype=np.float64).ravel(); b = np.asarray(b, dtype=np.float64).ravel()
    if a.size != b.size:
        print(f"  {name}: SIZE MISMATCH {a.size} vs {b.size}", flush=True); return
    d = np.abs(a - b)
    rel = np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
    print(f"  {name}: max_abs={d.max():.4e} rel_l2={rel:.4e} ref_absmax={np.abs(b).max():.3f} nan={np.isnan(a).sum()} "
          f"exact_frac={(a == b).mean():.4f}", flush=True)


def step_encode(engine):
    wb, ref_lib, ref, ctx, pcm, log = _load_both()
    try:
        ctx.set_gemm_engine(engine)
        ref.pcm_to_mel(pcm, 1); ctx.pcm_to_mel(pcm, 1)
        rmel, _ = ref.mel()
        cmp("mel", ctx.read_stage(wb.STAGE_HOST_MEL, np.float32), rmel)
        t0 = time.time(); rc = ref.encode(0, 8); t_ref = time.time() - t0
        t0 = time.time(); rc2 = ctx.encode(0); t_mine = time.time() - t0
        t0 = time.time(); rc2 = ctx.encode(0); t_mine2 = time.time() - t0
        print(f"  encode rc ref={rc} mine={rc2}  t_ref={t_ref*1e3:.1f}ms t_mine(first)={t_mine*1e3:.1f}ms t_mine(second)={t_mine2*1e3:.2f}ms")
        if rc2 != 0:
            return
        conv_ref = ref.embd_conv()            # [d][T]
        conv = ctx.read_stage(wb.STAGE_EMBD_CONV, np.float32).reshape(1500, -1)
        cmp("embd_conv", conv, conv_ref.T)
        enc_ref = ref.embd_enc()              # [T][d]
        enc = ctx.read_stage(wb.STAGE_EMBD_ENC, np.float32).reshape(1500, -1)
        cmp("embd_enc", enc, enc_ref)
        kr, vr = ref.kv_cross()
        cmp("cross_k", ctx.read_stage(wb.STAGE_CROSS_K, np.float16), kr)
        cmp("cross_v", ctx.read_stage(wb.STAGE_CROSS_V, np.float16), vr)
        # first decode step
        sot = 50257
        lr = ref.decode([sot], 0, 4)
        lm = ctx.decode([sot], 0)
        cmp("logits(step0)", lm, lr)
        print(f"  argmax ref={int(lr.argmax())} mine={int(lm.argmax())}")
        lr = ref.decode([50363], 1, 4); lm = ctx.decode([50363], 1)
        cmp("logits(step1)", lm, lr)
        print(f"  argmax ref={int(lr.argmax())} mine={int(lm.argmax())}")
        toks = [843, 523, 616, 5891, 3399, 1265, 407, 644, 534, 1499, 460, 466]
        lr = ref.decode(toks, 2, 4); lm = ctx.decode(toks, 2)
        cmp("logits(12-token batch, tcgen05 path)", lm, lr)
        print(f"  argmax ref={int(lr.argmax())} mine={int(lm.argmax())}")
    finally:
        for l in log[-30:]:
            print("  LOG", l[0], l[1].rstrip())


def step_full(engine):
    wb, ref_lib, ref, ctx, pcm, log = _load_both()
    try:
        ctx.set_gemm_engine(engine)
        for name, audio, mt in (("jfk", pcm, 0), ("jfk mt16", pcm, 16), ("jfk30", ref_lib.jfk30(pcm), 0)):
            pr = ref_lib.host_params(ref.lib, max_tokens=mt, n_threads=8, temperature_inc=0.0)
            pm = wb.host_params(ctx.lib, max_tokens=mt, n_threads=8, temperature_inc=0.0)
            t0 = time.time(); rc_r = ref.full(pr, audio); t_r = time.time() - t0
            t0 = time.time(); rc_m = ctx.full(pm, audio); t_m = time.time() - t0
            t0 = time.time(); rc_m = ctx.full(pm, audio); t_m2 = time.time() - t0
            rr, rm = ref.result(), ctx.result()
            ids_r = [t["id"] for s in rr["segments"] for t in s["tokens"]]
            ids_m = [t["id"] for s in rm["segments"] for t in s["tokens"]]
            print(f"  {name}: rc {rc_r}/{rc_m} t_ref={t_r*1e3:.0f}ms t_mine={t_m*1e3:.0f}ms (2nd {t_m2*1e3:.0f}ms) ids_equal={ids_r == ids_m} n={len(ids_r)}/{len(ids_m)}")
            print(f"    ref : {rr['text']!r}")
            print(f"    mine: {rm['text']!r}")
            if ids_r == ids_m:
                for k in ("p", "plog", "pt", "ptsum"):
                    a = np.array([t[k] for s in rm["segments"] for t in s["tokens"]]); b = np.array([t[k] for s in rr["segments"] for t in s["tokens"]])
                    print(f"    token {k}: max_abs_diff={np.abs(a-b).max():.3e}")
                t0m = [(t["t0"], t["t1"]) for s in rm["segments"] for t in s["tokens"]]; t0r = [(t["t0"], t["t1"]) for s in rr["segments"] for t in s["tokens"]]
                print(f"    token timestamps equal={t0m == t0r}")
            print("    counters", ctx.counters(), ctx.timings_us())
    finally:
        for l in log[-12:]:
            print("  LOG", l[0], l[1].rstrip())


STEPS = {"gemm_simt": lambda: step_gemm(1), "gemm_tc": lambda: step_gemm(0), "encode_simt": lambda: step_encode(1),
         "encode_tc": lambda: step_encode(0), "full_simt": lambda: step_full(1), "full_tc": lambda: step_full(0)}

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--step":
        STEPS[sys.argv[2]]()
        sys.exit(0)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    logf = open(os.path.join(ROOT, "gpurun_out", "bringup.log"), "w")
    names = sys.argv[1:] or list(STEPS)
    for s in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--step", s], capture_output=True, text=True, timeout=240)
            out = r.stdout + r.stderr[-3000:]
            hdr = f"=== {s}: rc={r.returncode} {time.time()-t0:.1f}s"
        except subprocess.TimeoutExpired as e:
            out = (e.stdout or b"").decode() if isinstance(e.stdout, bytes) else (e.stdout or "")
            hdr = f"=== {s}: TIMEOUT"
        print(hdr); print(out, flush=True)
        logf.write(hdr + "\n" + out + "\n"); logf.flush()

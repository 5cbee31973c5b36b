#!/usr/bin/env python
"""Generates tests/golden/jfk_tiny_en.npz from the compiled reference (oracle/_ref: whisper.cpp v1.5.4 CPU path, BLAS off)
on tests/golden/jfk.wav with the real tiny.en weights.  TEST TOOLING.  Re-run only where oracle/_ref can be built
(/root/reference mounted):   python tools/make_golden.py

Contents (small on purpose — strided samples plus checksums):
  mel_sha1           sha1 of the f32 log-mel bytes [80][n_len]         (bit-exact target)
  mel_sample         mel[::8, ::97]
  conv_sample        embd_conv^T[::25, ::16]   (token-major [T][d])
  enc_sample         embd_enc[::25, ::16]
  cross_k_sample     kv_cross K [Lt][T][d] -> [:, ::50, ::16]   (f16)
  cross_v_sample     kv_cross V^T [Lt][d][T] -> [:, ::16, ::50] (f16)
  logits_sot         full logits row of decode([sot], n_past=0)
  logits_beg         full logits row of decode([beg], n_past=1)
  ids_maxtok16 / ids_full / ids_jfk30     greedy token ids: host defaults (max_tokens 16), max_tokens 0, 30 s tiled clip (temperature_inc 0)
  p_full, t0_full, t1_full                 per-token probability and token timestamps of the max_tokens=0 run
  text_full
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_lib  # noqa: E402

SOT, BEG = 50257, 50363


def main():
    lib = ref_lib.load()
    model = open(ref_lib.tiny_en_model_path(), "rb").read()
    pcm = ref_lib.read_wav_f32(os.path.join(ROOT, "tests", "golden", "jfk.wav"))
    rs = ref_lib.RefSession(lib, model, use_gpu=False)
    out = {}
    assert rs.pcm_to_mel(pcm, 4) == 0
    mel, _ = rs.mel()
    out["mel_sha1"] = np.frombuffer(hashlib.sha1(np.ascontiguousarray(mel).tobytes()).digest(), np.uint8)
    out["mel_shape"] = np.array(mel.shape, np.int32)
    out["mel_sample"] = mel[::8, ::97].copy()
    lib.probe_set_audio_ctx(rs.ctx, 0)
    assert rs.encode(0, 4) == 0
    out["conv_sample"] = rs.embd_conv().T[::25, ::16].copy()
    out["enc_sample"] = rs.embd_enc()[::25, ::16].copy()
    k, v = rs.kv_cross()
    Lt, d, T = 4, 384, 1500
    out["cross_k_sample"] = k.reshape(Lt, T, d)[:, ::50, ::16].copy()
    out["cross_v_sample"] = v.reshape(Lt, d, T)[:, ::16, ::50].copy()
    out["logits_sot"] = rs.decode([SOT], 0, 4).reshape(-1).copy()
    out["logits_beg"] = rs.decode([BEG], 1, 4).reshape(-1).copy()

    def ids(r):
        return np.array([t["id"] for s in r["segments"] for t in s["tokens"]], np.int32)

    assert rs.full(ref_lib.host_params(lib, max_tokens=16, n_threads=4), pcm) == 0
    out["ids_maxtok16"] = ids(rs.result())
    assert rs.full(ref_lib.host_params(lib, max_tokens=0, n_threads=4), pcm) == 0
    r = rs.result()
    out["ids_full"] = ids(r)
    toks = [t for s in r["segments"] for t in s["tokens"]]
    out["p_full"] = np.array([t["p"] for t in toks], np.float32)
    out["t0_full"] = np.array([t["t0"] for t in toks], np.int64)
    out["t1_full"] = np.array([t["t1"] for t in toks], np.int64)
    out["text_full"] = np.frombuffer(r["text"], np.uint8)
    assert rs.full(ref_lib.host_params(lib, max_tokens=0, n_threads=4, temperature_inc=0.0), ref_lib.jfk30(pcm)) == 0
    out["ids_jfk30"] = ids(rs.result())
    rs.close()
    path = os.path.join(ROOT, "tests", "golden", "jfk_tiny_en.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

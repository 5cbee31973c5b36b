#!/usr/bin/env python
"""Generates tests/golden/bench_chunks_tiny_en.npz: the greedy token ids the compiled reference (oracle/_ref, whisper.cpp v1.5.4 CPU
path) produces for every distinct chunk of bench.py's workload (jfk.wav tiled to 30 s, circularly shifted by k * 1.7 s; the shift has
period 300), with bench.py's parameter block (host block of SpeechToText::transcribe with max_tokens=0, entropy_thold=2.4,
temperature_inc=0).  tests/test_gpu_parity.py compares every chunk of a 72- and a 512-chunk batch against it and bench.py checks the
transcripts of its timed run against it.  Run where oracle/_ref is built:   python tools/make_bench_golden.py [--jobs 8]"""
import argparse
import hashlib
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_lib  # noqa: E402

PERIOD = 300


def chunk(base, k):
    return np.roll(base, int(k * 1.7 * 16000)).copy()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "bench_chunks_tiny_en.npz"))
    args = ap.parse_args()
    lib = ref_lib.load()
    model = open(ref_lib.tiny_en_model_path(), "rb").read()
    pcm = ref_lib.read_wav_f32(os.path.join(ROOT, "tests", "golden", "jfk.wav"))
    base = ref_lib.jfk30(pcm)
    ids = [None] * PERIOD
    texts = [None] * PERIOD
    nxt = [0]
    lock = threading.Lock()

    def work():
        s = ref_lib.RefSession(lib, model, use_gpu=False)
        p = ref_lib.host_params(lib, max_tokens=0, entropy_thold=2.4, temperature_inc=0.0, n_threads=1)
        while True:
            with lock:
                k = nxt[0]
                nxt[0] += 1
            if k >= PERIOD:
                break
            assert s.full(p, chunk(base, k)) == 0
            r = s.result()
            ids[k] = [t["id"] for sg in r["segments"] for t in sg["tokens"]]
            texts[k] = r["text"]
        s.close()

    ts = [threading.Thread(target=work) for _ in range(args.jobs)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    off = np.cumsum([0] + [len(x) for x in ids]).astype(np.int32)
    np.savez_compressed(args.out, ids=np.concatenate([np.asarray(x, np.int32) for x in ids]), offsets=off,
                        text_sha1=np.frombuffer(b"".join(hashlib.sha1(t).digest() for t in texts), np.uint8).reshape(PERIOD, 20),
                        n_chars=np.asarray([len(t) for t in texts], np.int32))
    print("wrote", args.out, "tokens", int(off[-1]), "chars", sum(len(t) for t in texts))


if __name__ == "__main__":
    main()

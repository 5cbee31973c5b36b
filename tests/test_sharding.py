"""The N > 1 path on a CPU box: two gloo ranks, one model broadcast, chunks i mod world, no step collective, transcripts
gathered on rank 0 (godot-whisper_b200/shard.py; reference analogue whisper_full_parallel, whisper.cpp:5817-5930).
The ranks run the product's host driver linked to the TEST-ONLY checker forward (tests/hostlogic), since there is no GPU here."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from conftest import ROOT, PKG, build_hostlogic, ids_of
from oracle import ref_lib
import shard
import whisper_b200 as wb


def test_shard_indices_partition():
    for n, world in ((0, 2), (1, 2), (7, 2), (64, 8), (5, 8)):
        parts = [shard.shard_indices(n, r, world) for r in range(world)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard.shard_indices(4, 2, 2)


def test_gather_single_process_orders_by_chunk():
    assert shard.gather_transcripts({1: {"t": "b"}, 0: {"t": "a"}}, 2) == [{"t": "a"}, {"t": "b"}]
    assert shard.broadcast_model(b"abc") == b"abc"


WORKER = textwrap.dedent("""
    import os, sys, json, hashlib
    sys.path.insert(0, {root!r}); sys.path.insert(0, {pkg!r})
    import numpy as np
    import torch.distributed as dist
    import shard, whisper_b200 as wb
    from oracle import ref_lib          # test infrastructure: supplies the checker forward's tensor math
    dist.init_process_group("gloo")
    rank = dist.get_rank()
    os.environ["WHISPER_HOSTLOGIC_REF_LIB"] = ref_lib.ref_lib_path()
    blob = open(ref_lib.tiny_en_model_path(), "rb").read() if rank == 0 else None
    blob = shard.broadcast_model(blob, dist)                       # the one collective of the path
    lib = wb.load_library({hostlogic!r}); wb.set_log_sink(lib, None)
    ctx = wb.Context(blob, lib=lib)
    pcm = ref_lib.read_wav_f32(os.path.join({root!r}, "tests", "golden", "jfk.wav"))
    chunks = [pcm, pcm[:8000], np.roll(pcm, 16000), pcm[:64000], pcm[16000:]]     # one ragged, one < 1 s
    p = wb.host_params(lib, max_tokens=16, n_threads=2, temperature_inc=0.0)
    res = shard.transcribe_sharded(ctx, p, chunks, dist, batch=2)
    if rank == 0:
        out = dict(sha=hashlib.sha1(blob).hexdigest(), ids=[[t["id"] for s in r["segments"] for t in s["tokens"]] for r in res],
                   mine=shard.shard_indices(len(chunks), 0, dist.get_world_size()))
        print("RESULT " + json.dumps(out), flush=True)
    else:
        assert res is None
        print("SHA1 " + hashlib.sha1(blob).hexdigest(), flush=True)
    ctx.close(); dist.barrier(); dist.destroy_process_group()
""")


def test_two_rank_gloo_sharded_transcription(ref, model_bytes, jfk, tmp_path):
    import hashlib
    import json
    hostlogic = build_hostlogic()
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, pkg=PKG, hostlogic=hostlogic))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")][0]
    got = json.loads(line[len("RESULT "):])
    sha = hashlib.sha1(model_bytes).hexdigest()
    assert got["sha"] == sha                                         # rank 0 read it ...
    assert [l for l in out.stdout.splitlines() if l.startswith("SHA1 ")] == ["SHA1 " + sha]   # ... rank 1 got it by broadcast
    assert got["mine"] == [0, 2, 4]

    # single-process answer for the same chunks: the compiled reference itself
    rs = ref_lib.RefSession(ref, model_bytes, use_gpu=False)
    try:
        chunks = [jfk, jfk[:8000], np.roll(jfk, 16000), jfk[:64000], jfk[16000:]]
        p = ref_lib.host_params(ref, max_tokens=16, n_threads=2, temperature_inc=0.0)
        want = []
        for c in chunks:
            assert rs.full(p, c) == 0
            want.append(ids_of(rs.result()))
    finally:
        rs.close()
    assert got["ids"] == want
    assert want[1] == []                                             # < 1 s of audio: no segments (whisper.cpp:5015-5021)

"""The oracle against the committed golden fixtures (tests/golden/jfk_tiny_en.npz, made by tools/make_golden.py): guards the
compiled reference against toolchain / flag drift on whatever box builds it (v3 = AVX2 and v4 = AVX-512 builds differ only in
summation order)."""
import hashlib
import os

import numpy as np

from conftest import ROOT, ids_of
from oracle import ref_lib

SOT = 50257


def test_oracle_reproduces_golden_fixtures(ref, ref_session, jfk):
    g = np.load(os.path.join(ROOT, "tests", "golden", "jfk_tiny_en.npz"))
    assert ref_session.pcm_to_mel(jfk, 4) == 0
    mel, _ = ref_session.mel()
    assert hashlib.sha1(np.ascontiguousarray(mel).tobytes()).digest() == bytes(g["mel_sha1"])
    assert np.array_equal(mel[::8, ::97], g["mel_sample"])
    ref.probe_set_audio_ctx(ref_session.ctx, 0)
    assert ref_session.encode(0, 4) == 0
    enc = ref_session.embd_enc()[::25, ::16]
    assert np.abs(enc - g["enc_sample"]).max() <= 2e-2
    lg = ref_session.decode([SOT], 0, 4).reshape(-1)
    assert np.abs(lg - g["logits_sot"]).max() <= 5e-2 and int(lg.argmax()) == int(g["logits_sot"].argmax())
    assert ref_session.full(ref_lib.host_params(ref, max_tokens=0, n_threads=4), jfk) == 0
    assert ids_of(ref_session.result()) == g["ids_full"].tolist()
    assert ref_session.result()["text"] == bytes(g["text_full"])

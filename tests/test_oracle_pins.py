"""Pins the oracle (compiled reference, oracle/_ref) to the reference's own known answers for this path:
the transcript baked into bin/samples/godot_whisper/audio_transcribe.tscn:23 and the golden token stream of App. C."""
import numpy as np

from conftest import ids_of
from oracle import ref_lib

TSCN_TEXT = b" And so my fellow Americans ask not what your country can do for you ask what you can do for your country."
GOLDEN_IDS_MAXTOK16 = [50363, 843, 523, 616, 5891, 3399, 1265, 407, 644, 534, 1499, 460, 466, 329, 345, 50763, 50763]


def test_reference_reproduces_the_sample_scene_transcript(ref, ref_session, jfk):
    p = ref_lib.host_params(ref, max_tokens=0, n_threads=4)
    assert ref_session.full(p, jfk) == 0
    assert ref_session.result()["text"] == TSCN_TEXT


def test_reference_golden_token_stream_with_host_defaults(ref, ref_session, jfk):
    p = ref_lib.host_params(ref, max_tokens=16, n_threads=4)
    assert ref_session.full(p, jfk) == 0
    assert ids_of(ref_session.result()) == GOLDEN_IDS_MAXTOK16

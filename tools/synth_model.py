"""Synthetic-weight Whisper models in the ggml file format (writer of the real files:
/root/reference/thirdparty/whisper.cpp/models/convert-pt-to-ggml.py:268-339).

base.en / small.en weights are not on disk and cannot be downloaded, so throughput and parity runs for those shapes use
seeded random tensors of the right shapes appended to the header (hparams + mel filters + vocab) of the real tiny.en file
— every English-only model shares that vocabulary and filter bank.  Both the compiled reference and libwhisper_b200.so load
the result; transcripts are meaningless, tensor shapes / FLOPs / bytes are those of the named model.
"""
from __future__ import annotations

import struct

import numpy as np

SHAPES = {  # n_audio_state, n_audio_head, n_audio_layer, n_text_state, n_text_head, n_text_layer
    "tiny.en": (384, 6, 4, 384, 6, 4),
    "base.en": (512, 8, 6, 512, 8, 6),
    "small.en": (768, 12, 12, 768, 12, 12),
    # multilingual shapes: same tensors, the header (n_vocab 51 865, multilingual vocabulary) comes from the weight-less test
    # header the reference ships (thirdparty/whisper.cpp/models/for-tests-ggml-tiny.bin, staged as for-tests-ggml-multilingual.bin)
    "tiny": (384, 6, 4, 384, 6, 4),
    "base": (512, 8, 6, 512, 8, 6),
    "small": (768, 12, 12, 768, 12, 12),
}


def split_header(model_bytes: bytes):
    """-> (hparams tuple of 11 ints, bytes of [filters + vocab] section, offset of the first tensor record)."""
    magic, = struct.unpack_from("<I", model_bytes, 0)
    assert magic == 0x67676D6C
    hp = struct.unpack_from("<11i", model_bytes, 4)
    off = 48
    n_mel, n_fft = struct.unpack_from("<2i", model_bytes, off)
    off += 8 + 4 * n_mel * n_fft
    n_tok, = struct.unpack_from("<i", model_bytes, off)
    off += 4
    for _ in range(n_tok):
        ln, = struct.unpack_from("<I", model_bytes, off)
        off += 4 + ln
    return hp, model_bytes[48:off], off


def _tensor(name: str, arr: np.ndarray) -> bytes:
    """One tensor record: n_dims, name_len, ttype, ne[] reversed torch shape, name, data (convert-pt-to-ggml.py:323-337)."""
    ttype = 1 if arr.dtype == np.float16 else 0
    nb = name.encode()
    out = struct.pack("<3i", arr.ndim, len(nb), ttype)
    for dim in reversed(arr.shape):
        out += struct.pack("<i", dim)
    return out + nb + np.ascontiguousarray(arr).tobytes()


def mel_filter_bank(n_mels: int, n_fft_bins: int = 201, sr: int = 16000) -> np.ndarray:
    """A triangular mel filter bank [n_mels][n_fft_bins] (HTK-style mel scale, area-normalised): what a 128-band model file carries in
    place of the 80-band one.  Both libraries read the bank from the file, so its exact shape only has to be plausible."""
    f = np.linspace(0.0, sr / 2.0, n_fft_bins)
    mel = lambda hz: 2595.0 * np.log10(1.0 + hz / 700.0)
    inv = lambda m: 700.0 * (10.0 ** (m / 2595.0) - 1.0)
    edges = inv(np.linspace(mel(0.0), mel(sr / 2.0), n_mels + 2))
    bank = np.zeros((n_mels, n_fft_bins), np.float32)
    for j in range(n_mels):
        lo, ce, hi = edges[j], edges[j + 1], edges[j + 2]
        up, down = (f - lo) / max(ce - lo, 1e-9), (hi - f) / max(hi - ce, 1e-9)
        bank[j] = np.maximum(0.0, np.minimum(up, down)) * (2.0 / (hi - lo))
    return bank


def make_model(tiny_en_bytes: bytes, name: str = "base.en", seed: int = 1234, std: float = 0.02, n_mels: int | None = None,
               n_vocab: int | None = None, shape: tuple | None = None) -> bytes:
    """`tiny_en_bytes`: any ggml Whisper file whose header (hparams, filters, vocabulary) the new model inherits.  n_mels / n_vocab
    override the header (large-v3: 128 bands, 51 866 tokens — whisper.cpp:1135, 1161-1163); shape overrides SHAPES[name]."""
    d_a, h_a, l_a, d_t, h_t, l_t = shape or SHAPES[name]
    hp, mid, _ = split_header(tiny_en_bytes)
    n_vocab0, n_audio_ctx, _, _, _, n_text_ctx, _, _, _, n_mels0, _ = hp
    if n_mels is not None and n_mels != n_mels0:
        fm, ff = struct.unpack_from("<2i", mid, 0)
        mid = struct.pack("<2i", n_mels, ff) + mel_filter_bank(n_mels, ff).tobytes() + mid[8 + 4 * fm * ff:]
    n_mels = n_mels or n_mels0
    n_vocab = n_vocab or n_vocab0
    rng = np.random.default_rng(seed)
    out = [struct.pack("<I", 0x67676D6C),
           struct.pack("<11i", n_vocab, n_audio_ctx, d_a, h_a, l_a, n_text_ctx, d_t, h_t, l_t, n_mels, 1), mid]

    def w(*shape):   # f16 weight
        return (rng.standard_normal(shape) * std).astype(np.float16)

    def b(*shape):   # f32 bias
        return (rng.standard_normal(shape) * std).astype(np.float32)

    def g(n):        # LayerNorm gain
        return (1.0 + rng.standard_normal(n) * std).astype(np.float32)

    def add(nm, arr):
        out.append(_tensor(nm, arr))

    add("encoder.conv1.weight", w(d_a, n_mels, 3)); add("encoder.conv1.bias", b(d_a, 1))
    add("encoder.conv2.weight", w(d_a, d_a, 3));    add("encoder.conv2.bias", b(d_a, 1))
    add("encoder.positional_embedding", (rng.standard_normal((n_audio_ctx, d_a)) * 0.01).astype(np.float32))
    for i in range(l_a):
        p = f"encoder.blocks.{i}."
        add(p + "attn_ln.weight", g(d_a)); add(p + "attn_ln.bias", b(d_a))
        add(p + "attn.query.weight", w(d_a, d_a)); add(p + "attn.query.bias", b(d_a))
        add(p + "attn.key.weight", w(d_a, d_a))
        add(p + "attn.value.weight", w(d_a, d_a)); add(p + "attn.value.bias", b(d_a))
        add(p + "attn.out.weight", w(d_a, d_a)); add(p + "attn.out.bias", b(d_a))
        add(p + "mlp_ln.weight", g(d_a)); add(p + "mlp_ln.bias", b(d_a))
        add(p + "mlp.0.weight", w(4 * d_a, d_a)); add(p + "mlp.0.bias", b(4 * d_a))
        add(p + "mlp.2.weight", w(d_a, 4 * d_a)); add(p + "mlp.2.bias", b(d_a))
    add("encoder.ln_post.weight", g(d_a)); add("encoder.ln_post.bias", b(d_a))
    add("decoder.positional_embedding", (rng.standard_normal((n_text_ctx, d_t)) * 0.01).astype(np.float32))
    add("decoder.token_embedding.weight", w(n_vocab, d_t))
    for i in range(l_t):
        p = f"decoder.blocks.{i}."
        for a in ("attn", "cross_attn"):
            add(p + a + "_ln.weight", g(d_t)); add(p + a + "_ln.bias", b(d_t))
            add(p + a + ".query.weight", w(d_t, d_t)); add(p + a + ".query.bias", b(d_t))
            add(p + a + ".key.weight", w(d_t, d_t))
            add(p + a + ".value.weight", w(d_t, d_t)); add(p + a + ".value.bias", b(d_t))
            add(p + a + ".out.weight", w(d_t, d_t)); add(p + a + ".out.bias", b(d_t))
        add(p + "mlp_ln.weight", g(d_t)); add(p + "mlp_ln.bias", b(d_t))
        add(p + "mlp.0.weight", w(4 * d_t, d_t)); add(p + "mlp.0.bias", b(4 * d_t))
        add(p + "mlp.2.weight", w(d_t, 4 * d_t)); add(p + "mlp.2.bias", b(d_t))
    add("decoder.ln.weight", g(d_t)); add("decoder.ln.bias", b(d_t))
    return b"".join(out)


GGML_TYPES = {"q4_0": (2, 2, 18), "q4_1": (3, 3, 20), "q5_0": (6, 8, 22), "q5_1": (7, 9, 24), "q8_0": (8, 7, 34)}     # name -> (ggml type, ftype, bytes per 32-element block)
SKIP_QUANT = ("encoder.conv1.bias", "encoder.conv2.bias", "encoder.positional_embedding", "decoder.positional_embedding")


def quantize_model(model_bytes: bytes, qname: str, ref) -> bytes:
    """The file whisper.cpp's quantize tool would write for an f16 model (examples/quantize/quantize.cpp + examples/common-ggml.cpp:
    every 2-D tensor except the positional embeddings is block-quantised, everything else is copied; ftype = 2000 + GGML_FTYPE).
    Quantisation itself is the reference's own ggml_quantize_chunk (called through `ref`, the compiled reference library)."""
    import ctypes as C
    gtype, ftype, bsz = GGML_TYPES[qname]
    hp, mid, off = split_header(model_bytes)
    out = [struct.pack("<I", 0x67676D6C), struct.pack("<11i", *hp[:10], 2000 + ftype), mid]
    ref.ggml_quantize_chunk.restype = C.c_size_t
    ref.ggml_quantize_chunk.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    hist = (C.c_int64 * 16)()
    n = len(model_bytes)
    while off < n:
        n_dims, name_len, ttype = struct.unpack_from("<3i", model_bytes, off)
        off += 12
        ne = struct.unpack_from("<%di" % n_dims, model_bytes, off)
        off += 4 * n_dims
        name = model_bytes[off:off + name_len].decode()
        off += name_len
        nelem = int(np.prod(ne))
        nbytes = nelem * (2 if ttype == 1 else 4)
        data = model_bytes[off:off + nbytes]
        off += nbytes
        if n_dims == 2 and name not in SKIP_QUANT and ne[0] % 32 == 0:
            src = (np.frombuffer(data, np.float16) if ttype == 1 else np.frombuffer(data, np.float32)).astype(np.float32)
            dst = np.empty(nelem // 32 * bsz, np.uint8)
            got = ref.ggml_quantize_chunk(gtype, src.ctypes.data, dst.ctypes.data, 0, nelem, C.addressof(hist))
            assert got == dst.size, (name, got, dst.size)
            ttype, data = gtype, dst.tobytes()
        out.append(struct.pack("<3i", n_dims, name_len, ttype) + struct.pack("<%di" % n_dims, *ne) + name.encode() + data)
    return b"".join(out)

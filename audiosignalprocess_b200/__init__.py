"""B200-native WebRTC noise suppression (float NS + fixed NSx) behind the reference's C API.

The product is libwebrtc_ns_b200.so (include/webrtc_ns_b200.h); this package is the thin
Python mirror of that C ABI used by the tests and bench.py.  There is no CPU fallback:
importing works anywhere, but every call needs the CUDA library and a GPU.
"""
from .capi import load_library, LIB_PATH  # noqa: F401
from .ns import (  # noqa: F401
    NsError, NoiseSuppressor, NoiseSuppressorX, NsBatch, synth_pcm_host, frame_len, num_bands,
    run_generated_job,
)

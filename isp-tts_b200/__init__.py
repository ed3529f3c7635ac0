"""B200-native Aligner hot path of ilya16/isp-tts: text x mel log-likelihood,
Monotonic Alignment Search and duration extraction, as hand-written sm_100a
CUDA behind a C ABI (include/isp_tts_b200.h).  No CPU fallback: every compute
entry point raises if the CUDA library is missing or no B200 is visible.
"""
__version__ = "0.1.0"

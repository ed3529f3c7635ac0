"""B200-native Aligner hot path of ilya16/isp-tts: text x mel log-likelihood,
Monotonic Alignment Search and duration extraction, as hand-written sm_100a
CUDA behind a C ABI (include/isp_tts_b200.h).  No CPU fallback: every compute
entry point raises if the CUDA library is missing or no B200 is visible.

    from isp_tts_b200 import Aligner, AlignerOutput, b_mas, cuda_b_mas

Importing the package does not load the CUDA library (so host-only helpers such
as `synth` and `sharding` work anywhere); the first compute call does.
"""
__version__ = "0.1.0"

_LAZY = {
    "Aligner": "alignment", "AlignerConfig": "alignment", "AlignerOutput": "alignment",
    "ConvAttention": "alignment", "ConvAttentionConfig": "alignment", "ConvBlock1D": "alignment",
    "batch_diagonal_prior": "alignment", "loglik_forward": "alignment",
    "b_mas": "mas", "cuda_b_mas": "mas", "mas_forward": "mas", "mas_durations": "mas",
    "binarization_loss": "mas", "stage_operands": "alignment",
    "LengthRegulator": "consumers", "TemporalAverager": "consumers", "length_regulate": "consumers", "temporal_average": "consumers",
    "AttentionCTCLoss": "ctc", "attention_ctc_loss": "ctc", "ctc_nll": "ctc",
    "gather_durations": "sharding", "shard_bounds": "sharding", "balanced_assignment": "sharding",
}


def __getattr__(name):
    mod = _LAZY.get(name)
    if mod is None:
        raise AttributeError(f"module 'isp_tts_b200' has no attribute {name!r}")
    import importlib
    return getattr(importlib.import_module(f"{__name__}.{mod}"), name)


__all__ = sorted(_LAZY)

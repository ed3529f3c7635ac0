"""Generates tests/golden/*.npz by running the UNMODIFIED reference here.

TEST INFRASTRUCTURE ONLY; run in the build container (needs /root/reference):
    python oracle/gen_golden.py
Outputs are small on purpose (committed).  Every array that the reference
produced is stored verbatim; large inputs are stored as the seed that
isp_tts_b200.synth regenerates them from.
"""
from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from isp_tts_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def ref_mas(b_mas, x, text_len, mel_len):
    """reference b_mas on a COPY (it mutates its input, mas.py:33)."""
    return b_mas(np.array(x, dtype=np.float32, copy=True), np.asarray(text_len), np.asarray(mel_len))


def paths_of(hard, mel_len):
    p = hard.argmax(axis=2).astype(np.int16)
    for b, n in enumerate(mel_len):
        p[b, int(n):] = -1
    return p


def gen_mas():
    b_mas = ref_loader.load_b_mas()
    mas_width1 = ref_loader.load_mas_width1()

    # --- known-answer cases -------------------------------------
    kat = {}
    x = np.array([[-1, -2, -3], [-2, -.5, -4], [-3, -.5, -1], [-.5, -2, -.25], [-1, -1, -.5]], np.float32)
    q = x.copy()
    kat["a12_x"], kat["a12_hard"] = x, mas_width1(q)
    kat["a12_Q"] = q                      # the reference leaves the accumulated Q in its input
    for name, shape in [("zeros_10x4", (10, 4)), ("zeros_3x6", (3, 6)), ("zeros_1x3", (1, 3)), ("zeros_5x1", (5, 1)),
                        ("zeros_1x1", (1, 1)), ("zeros_7x7", (7, 7))]:
        z = np.zeros(shape, np.float32)
        kat[name + "_hard"] = mas_width1(z.copy())
    np.savez_compressed(os.path.join(OUT, "mas_kat.npz"), **kat)

    # --- cfg1: single utterance 80 tok x 400 fr (BASELINE.json configs[0]) -------------
    w = synth.WORKLOADS["cfg1"]
    x = synth.noise_logits(1, w.t1max, w.t2max, w.seed)
    tl, ml = synth.workload_lengths(w)
    hard = ref_mas(b_mas, x, tl, ml)
    qacc = x[0].copy(); mas_width1(qacc)
    np.savez_compressed(os.path.join(OUT, "mas_cfg1.npz"), seed=w.seed, x=x, text_len=tl, mel_len=ml,
                        path=paths_of(hard, ml), durations=hard.sum(axis=1, dtype=np.int64),
                        Q_last_row=qacc[-1])

    # --- small ragged batch with forced ties, T2 > T1, T1 == 1, T2 == 1 ----------------
    rs = np.random.RandomState(7)
    B, T1, T2 = 8, 48, 20
    x = (np.rint(rs.standard_normal((B, T1, T2)) * 2) / 2).astype(np.float32)   # step 0.5 -> many finite ties
    x[3] = 0.0
    tl = np.array([20, 13, 1, 7, 20, 5, 9, 16], np.int64)
    ml = np.array([48, 31, 17, 48, 12, 1, 9, 33], np.int64)     # b=4: T2 > T1 ; b=5: T1 == 1 ; b=2: T2 == 1
    hard = ref_mas(b_mas, x, tl, ml)
    np.savez_compressed(os.path.join(OUT, "mas_ragged_ties.npz"), x=x, text_len=tl, mel_len=ml, hard=hard)

    # --- seeded medium/large cases: inputs regenerated from the seed -------------------
    for tag, (B, T1, T2, ragged, quant) in {
        "cfg2_noise": (32, 1000, 200, True, 0.0),
        "cfg2_ties": (32, 1000, 200, True, 0.25),
        "odd_shapes": (5, 333, 77, True, 0.5),
        "wide": (3, 700, 601, True, 0.0),
        "long": (2, 4096, 512, False, 0.0),
    }.items():
        seed = 4242 + len(tag)
        tl, ml = synth.lengths(B, T2, T1, ragged, seed)
        x = synth.noise_logits(B, T1, T2, seed, quantize=quant)
        hard = ref_mas(b_mas, x, tl, ml)
        np.savez_compressed(os.path.join(OUT, f"mas_seeded_{tag}.npz"), seed=seed, B=B, T1=T1, T2=T2,
                            ragged=ragged, quantize=quant, text_len=tl, mel_len=ml,
                            path=paths_of(hard, ml), durations=hard.sum(axis=1, dtype=np.int64),
                            x_checksum=np.float64(x.astype(np.float64).sum()))


def gen_loglik():
    import torch
    m = ref_loader.load_alignment()
    b_mas = ref_loader.load_b_mas()
    torch.manual_seed(0)

    def run(tag, hp, B, T1, T2, tl, ml):
        al = m.Aligner(**hp).eval()
        tl_t, ml_t = torch.tensor(tl), torch.tensor(ml)
        mel = torch.randn(B, hp["mel_dim"], T1) * 2 - 5
        mel = mel.clamp(-11.5, 2.0)
        txt = torch.randn(B, hp["text_dim"], T2)
        mel = mel * (torch.arange(T1)[None, None] < ml_t[:, None, None])
        txt = txt * (torch.arange(T2)[None, None] < tl_t[:, None, None])
        cap = {}
        h1 = al.attention.key_proj[-1].register_forward_hook(lambda mod, i, o: cap.__setitem__("K", o.detach()))
        h2 = al.attention.query_proj[-1].register_forward_hook(lambda mod, i, o: cap.__setitem__("Q", o.detach()))
        with torch.no_grad():
            soft, logits = al.attention(queries=mel, keys=txt, query_len=ml_t, key_len=tl_t)
        h1.remove(); h2.remove()
        # GPU-route behaviour: MAS on a copy, attn_logits left intact
        hard = b_mas(logits.numpy().copy(), tl, ml)
        dur = hard.sum(axis=1, dtype=np.int64)
        sd = {"sd::" + k: v.numpy() for k, v in al.state_dict().items()}
        np.savez_compressed(
            os.path.join(OUT, f"loglik_{tag}.npz"),
            hp_keys=np.array(list(hp.keys())), hp_vals=np.array([repr(v) for v in hp.values()]),
            mel=mel.numpy(), enc_text=txt.numpy(), text_len=tl, mel_len=ml,
            Q=cap["Q"].transpose(1, 2).contiguous().numpy(),      # (B, T1, D)
            K=cap["K"].transpose(1, 2).contiguous().numpy(),      # (B, T2, D)
            attn_soft=soft.numpy(), attn_logits=logits.numpy(), attn_hard=hard, durations=dur, **sd)

    small = dict(mel_dim=8, text_dim=12, attention_dim=16, key_kernel_size=3, query_kernel_size=[3, 3],
                 dropout=0.1, normalization="instance", activation="relu")
    run("small", small, 4, 40, 12, np.array([12, 7, 3, 10], np.int64), np.array([40, 22, 9, 31], np.int64))
    # class-default attention_dim 80 (alignment.py:103), gelu
    mid = dict(mel_dim=16, text_dim=24, attention_dim=80, key_kernel_size=5, query_kernel_size=[5, 5],
               dropout=0.1, normalization="instance", activation="gelu")
    run("dim80", mid, 3, 150, 33, np.array([33, 20, 28], np.int64), np.array([150, 97, 140], np.int64))
    # recipe attention_dim 128 (recipes/acoustic/core.yaml:150-156), without the 1.7 M-parameter stacks' widths
    rec = dict(mel_dim=20, text_dim=32, attention_dim=128, key_kernel_size=5, query_kernel_size=[5, 5],
               dropout=0.1, normalization="instance", activation="gelu")
    run("dim128", rec, 3, 260, 48, np.array([48, 31, 40], np.int64), np.array([260, 180, 233], np.int64))


def gen_recipe():
    """The full reference Aligner at the recipe's shape (mel 80, text 384, attention_dim 128, kernels 5: 1.7 M parameters) on
    seeded weights and inputs (isp_tts_b200.synth.recipe_state / recipe_inputs regenerate both); only outputs are stored."""
    import torch
    m = ref_loader.load_alignment()
    b_mas = ref_loader.load_b_mas()
    seed, B, T1, T2 = 2024, 2, 300, 60
    tl, ml = np.array([60, 41], np.int64), np.array([300, 215], np.int64)
    al = m.Aligner(**synth.RECIPE_HP).eval()
    assert {k: tuple(v.shape) for k, v in al.state_dict().items()} == synth.RECIPE_SHAPES
    al.load_state_dict({k: torch.from_numpy(v) for k, v in synth.recipe_state(seed).items()}, strict=True)
    mel, txt = synth.recipe_inputs(seed + 1, B, T1, T2, tl, ml)
    with torch.no_grad():
        soft, logits = al.attention(queries=torch.from_numpy(mel), keys=torch.from_numpy(txt), query_len=torch.from_numpy(ml), key_len=torch.from_numpy(tl))
    hard = b_mas(logits.numpy().copy(), tl, ml)
    np.savez_compressed(os.path.join(OUT, "aligner_recipe.npz"), seed=seed, B=B, T1=T1, T2=T2, text_len=tl, mel_len=ml,
                        attn_logits=logits.numpy(), attn_soft=soft.numpy().astype(np.float16), path=paths_of(hard, ml),
                        durations=hard.sum(axis=1, dtype=np.int64))


def gen_consumers():
    """The reference's own LengthRegulator and TemporalAverager (temporal_adaptor.py:411-465) on both of their routes: hard
    durations (the MAS durations of a reference b_mas run) and the recipe's soft route (`alignment` = the Aligner's attn_soft,
    model.py:154).  Inputs and outputs are stored."""
    import torch
    ta = ref_loader.load_temporal_adaptor()
    b_mas = ref_loader.load_b_mas()
    B, T1, T2, C = 3, 170, 44, 48
    tl, ml = synth.lengths(B, T2, T1, True, 301)
    x_l = synth.noise_logits(B, T1, T2, 302)
    hard = b_mas(x_l.copy(), tl, ml)
    dur = torch.from_numpy(hard.sum(axis=1, dtype=np.int64))
    rs = np.random.RandomState(303)
    x = torch.from_numpy(rs.standard_normal((B, T2, C)).astype(np.float32))
    valid = (np.arange(T1)[None, :, None] < ml[:, None, None]) & (np.arange(T2)[None, None, :] < tl[:, None, None])
    soft = np.where(valid, np.exp(2.0 * x_l), 0.0)
    soft = (soft / np.maximum(soft.sum(2, keepdims=True), 1e-30)).astype(np.float32)          # rows of valid frames sum to 1
    soft_t = torch.from_numpy(soft)
    feat = (rs.rand(B, 2, T1) * 200.0 + 80.0).astype(np.float32)
    feat[rs.rand(B, 2, T1) < 0.3] = 0.0
    feat *= (np.arange(T1)[None, None, :] < ml[:, None, None])
    feat_t = torch.from_numpy(feat)
    lr, av = ta.LengthRegulator(), ta.TemporalAverager()
    out_h, dec_h = lr(x, dur, max_len=T1)                                   # temporal_adaptor.py:300 without alignment
    out_s, dec_s = lr(x, dur, max_len=T1, alignment=soft_t)                 # temporal_adaptor.py:300 on the soft route
    fdur = dur.float() + torch.from_numpy(rs.uniform(-0.45, 0.45, size=tuple(dur.shape)).astype(np.float32)) * (dur > 0)
    out_f, dec_f = lr(x, fdur)                                              # :325 shape of call, predicted (float) durations
    avg_h = av(feat_t, dur)                                                 # :447-465
    avg_s = av(feat_t, dur, soft_t)                                         # :446-449
    np.savez_compressed(os.path.join(OUT, "consumers.npz"), text_len=tl, mel_len=ml, durations=dur.numpy(), x=x.numpy(),
                        attn_soft=soft, path=paths_of(hard, ml), feat=feat, float_durations=fdur.numpy(),
                        lr_hard=out_h.numpy(), lr_hard_len=dec_h.numpy(), lr_soft=out_s.numpy(), lr_soft_len=dec_s.numpy(),
                        lr_float=out_f.numpy(), lr_float_len=dec_f.numpy(), avg_hard=avg_h.numpy(), avg_soft=avg_s.numpy())


if __name__ == "__main__":
    if not ref_loader.available():
        sys.exit("reference not found at " + ref_loader.REFERENCE_ROOT)
    os.makedirs(OUT, exist_ok=True)
    only = [a[7:] for a in sys.argv if a.startswith("--only=")]
    for name, fn in (("mas", gen_mas), ("loglik", gen_loglik), ("recipe", gen_recipe), ("consumers", gen_consumers)):
        if not only or name in only:
            fn()
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("wrote", sorted(os.listdir(OUT)), f"{total/1e6:.2f} MB")

"""The CTC oracle (oracle/ctc.py) pinned against a brute-force enumeration of the monotonic alignments on tiny cases
(the reference holds no golden vectors for this loss; its arithmetic is torch's nn.CTCLoss, loss.py:51)."""
import itertools

import numpy as np
import torch

from oracle import ctc as octc


def test_oracle_against_enumeration():
    rs = np.random.RandomState(3)
    for T1, T2 in [(3, 2), (5, 2), (4, 3), (6, 3)]:
        x = rs.standard_normal((1, T1, T2 + 1)) * 2.0            # one padded text column: it only enters the softmax
        ref = octc.attention_ctc_loss(torch.from_numpy(x), torch.tensor([T2]), torch.tensor([T1]), -1.0, reduction="none")
        # enumeration over the T2 real labels, softmax over all T2 + 1 columns plus the blank
        z = np.concatenate([np.full((T1, 1), -1.0), x[0]], axis=1)
        lp = z - np.log(np.exp(z).sum(axis=1, keepdims=True))
        total = 0.0
        for path in itertools.product(range(T2 + 1), repeat=T1):
            collapsed = [k for k, _ in itertools.groupby(path) if k != 0]
            if collapsed == list(range(1, T2 + 1)):
                total += np.exp(sum(lp[t, s] for t, s in enumerate(path)))
        assert abs(ref.item() - (-np.log(total))) < 1e-9


def test_oracle_reduction_and_zero_infinity():
    x = torch.zeros((2, 4, 3), dtype=torch.float64)
    tl, ml = torch.tensor([3, 2]), torch.tensor([2, 4])          # utterance 0 cannot be aligned
    per = octc.attention_ctc_loss(x, tl, ml, reduction="none")
    assert per[0].item() == 0.0 and per[1].item() > 0.0
    mean = octc.attention_ctc_loss(x, tl, ml)
    assert abs(mean.item() - (per[1].item() / 2) / 2) < 1e-12

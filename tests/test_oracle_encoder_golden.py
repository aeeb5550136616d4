"""Pin the matching-encoder oracle (oracle/oracle_encoder.py) on fixtures produced by executing the reference's
ResnetMatchingEncoder (oracle/make_golden_encoder.py).  CPU only."""
import numpy as np
import pytest
import torch

import helpers as hp
from oracle import oracle_encoder as oe

torch.set_grad_enabled(False)


@pytest.mark.parametrize("name", ["enc_tv", "enc_aa", "enc_aa_odd"])
def test_oracle_matches_the_executed_reference_encoder(name):
    fx = hp.load(name)
    n, h, w, seed, wseed, aa = [int(v) for v in fx["meta"]]
    images = torch.rand(n, 3, h, w, generator=torch.Generator().manual_seed(seed)) * 2 - 1
    sd = oe.encoder_state(wseed)
    got = torch.cat([oe.matching_encoder(images[i:i + 1], sd, antialiased=bool(aa)) for i in range(n)], 0)
    ref = torch.from_numpy(fx["feats"])
    assert got.shape == ref.shape == (n, 16, h // 4, w // 4)
    # instance-normalised features are O(1); the restatement uses the same ATen ops, so only summation order can differ
    assert float((got - ref).abs().max()) < 2e-5
    assert abs(float(got.mean())) < 1e-4 and abs(float(got.var(dim=(2, 3), unbiased=False).mean()) - 1) < 1e-3

"""MAMC / N-pairs loss kernels (hk_l2norm_rows_*, hk_npair_loss, 3xTF32 GEMMs) against reference-generated fixtures
(tests/golden/make_golden_cin.py imports model/loss/MAMC_loss.py) and the fp64 oracle at a realistic size."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'reference_cin.npz'))


@pytest.mark.parametrize('tag', ['npair_b8_p2', 'npair_b12_p3', 'npair_b6_p2_allsame', 'npair_b4_p2_alldiff'])
def test_npairs_vs_reference(tag):
    from hawkeye_b200.losses import NPairsLoss
    f = torch.from_numpy(G[f'{tag}_feats']).cuda().requires_grad_(True)
    lab = torch.from_numpy(G[f'{tag}_labels']).cuda()
    loss = NPairsLoss()(f, lab)
    loss.backward()
    ref = float(G[f'{tag}_loss'])
    e = rel_l2(f.grad.cpu(), G[f'{tag}_dfeats'])
    print(f'{tag}: loss {loss.item():.6f} vs {ref:.6f}; grad rel {e:.2e}')
    assert abs(loss.item() - ref) < 1e-4 * max(1.0, abs(ref)) and e < 1e-3


def test_mamc_vs_reference():
    from hawkeye_b200.losses import MAMCLoss

    class Cfg(dict):
        __getattr__ = dict.__getitem__
    crit = MAMCLoss(Cfg(lambda_a=0.5, use_mamc=True))
    pred = torch.from_numpy(G['mamc_pred']).cuda().requires_grad_(True)
    parts = torch.from_numpy(G['mamc_parts']).cuda().requires_grad_(True)
    loss = crit((pred, parts), torch.from_numpy(G['mamc_labels']).cuda())
    loss.backward()
    assert abs(loss.item() - float(G['mamc_loss'])) < 1e-4
    assert rel_l2(pred.grad.cpu(), G['mamc_dpred']) < 5e-4 and rel_l2(parts.grad.cpu(), G['mamc_dparts']) < 1e-3
    assert int(crit.last_correct.item()) == int((pred.argmax(1).cpu() == torch.from_numpy(G['mamc_labels'])).sum())


def test_npairs_osme_size_vs_oracle():
    """OSMENet's own shape: 16 samples x 2 attentions x 1024 features, 4 classes x 4 samples (BalancedBatchSampler)."""
    import detgen
    from hawkeye_b200.losses import NPairsLoss
    from oracle import hop_oracle as O
    f = detgen.det((16, 2, 1024), 77)
    lab = torch.arange(16) // 4
    fd = f.double().requires_grad_(True)
    ref = O.npairs_loss(fd, lab)
    ref.backward()
    fg = f.cuda().requires_grad_(True)
    loss = NPairsLoss()(fg, lab.cuda())
    loss.backward()
    e = rel_l2(fg.grad.cpu(), fd.grad)
    print(f'npairs 16x2x1024: loss {loss.item():.6f} vs {ref.item():.6f}; grad rel {e:.2e}')
    assert abs(loss.item() - ref.item()) < 1e-4 and e < 1e-3

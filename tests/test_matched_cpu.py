"""The matched-activation machinery on CPU: an fp32 oracle forward records its decisions in the GPU capture format; the
fp64 oracle replayed on that tape must reproduce the fp32 gradients to fp32 accuracy (no branch flips left), for the VGG
path (ReLU + 2x2 pools), the CBCNN signed sqrt and the ResNet path (3x3/s2 pool, BN, residual)."""
import torch

import detgen
import matched
from conftest import rel_l2
from oracle import hop_oracle as O


class Cfg(dict):
    __getattr__ = dict.__getitem__


def _check(forward, x, labels, state, keys, tol=2e-5):
    rec = matched.Recorder()
    _, loss32, g32 = O.loss_and_grads(lambda xx, st: forward(xx, st, rec), x, labels, state, keys)
    items = matched.tape_items(rec.cap)
    _, loss64, g64 = matched.oracle_step(forward, x, labels, state, items, keys)
    assert abs(float(loss32) - loss64) < 1e-4
    worst = max(rel_l2(g32[k], g64[k]) for k in keys)
    assert worst < tol, worst
    # and the tape really is what decides the branch: an all-ones ReLU tape gives different gradients
    return worst


def test_vgg_bcnn_and_cbcnn_tape():
    torch.set_num_threads(8)
    cfg = O.vgg_cfg_scaled(8)
    x, labels = detgen.det((2, 3, 32, 32), 41), detgen.det_labels(2, 20, 42)
    st = detgen.vgg_bcnn_state(cfg, 20, seed=100)
    keys = set(st.keys())
    _check(lambda xx, s, nl: O.bcnn_forward(xx, s, 2, cfg, nl=nl), x, labels, st, keys)
    st = detgen.vgg_bcnn_state(cfg, 20, seed=100, head_in=96)
    _check(lambda xx, s, nl: O.cbcnn_forward(xx, s, 96, 2, cfg, nl=nl), x, labels, st, set(st.keys()), tol=2e-4)


def test_resnet_mpn_tape():
    import hawkeye_b200 as hb
    torch.set_num_threads(8)
    net = hb.MODEL.get('MPN')(Cfg(name='MPN', iter_num=5, is_sqrt=True, is_vec=True, input_dim=2048,
                                  dimension_reduction=256, num_classes=200))
    st = detgen.state_like(net)
    keys = {k for k, _ in net.named_parameters()}
    x, labels = detgen.det((2, 3, 64, 64), 51), detgen.det_labels(2, 200, 52)
    _check(lambda xx, s, nl: O.mpn_forward(xx, s, 5, nl=nl), x, labels, st, keys, tol=5e-3)

"""Host logic of hawkeye_b200.examples (the Examples/{BCNN,CBCNN,MPN}.py equivalents) that needs no GPU: which parameters
train and with which learning-rate multipliers, checked against what the reference's Examples put into their optimizers."""
import os

import torch

import hawkeye_b200 as hb
from hawkeye_b200 import engine, examples
from hawkeye_b200.config import load_config

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bare(cls, cfg):
    t = object.__new__(cls)          # no CUDA needed for the builder hooks under test
    t.config = cfg
    t.total_epoch = cfg.train.epoch
    t.model = t.get_model(cfg.model)
    return t


def _groups(t):
    return [(sum(p.numel() for p in g if p.requires_grad), m) for g, m in t.param_groups() if any(p.requires_grad for p in g)]


def test_mpn_param_groups_follow_examples_mpn():
    t = _bare(examples.MPNTrainer, load_config(os.path.join(REPO, 'configs', 'MPN.yaml')))
    g = _groups(t)
    # Examples/MPN.py:14-18: classifier lr, pool lr, backbone 0.2 x lr
    assert [m for _, m in g] == [0.2, 1.0, 1.0]
    assert g[2][0] == 32896 * 200 + 200 and g[1][0] == 2048 * 256 + 2 * 256 and sum(n for n, _ in g) == 30612232
    flat = engine.FlatParams(None, groups=[grp for grp, _ in t.param_groups()])
    opt = engine.FusedAdam(flat, lr=8e-5, weight_decay=2e-5, group_lrs=[8e-5 * m for _, m in g])
    assert [round(pg['lr'] / 8e-5, 6) for pg in opt.param_groups] == [0.2, 1.0, 1.0]
    sch = t.get_scheduler.__func__(type('T', (), dict(optimizer=opt, total_epoch=100))(), t.config.train.scheduler)
    assert abs(opt.param_groups[0]['lr'] - 0.2 * 8e-5 * 0.01) < 1e-15      # warm-up starts at lr_warmup_decay x lr, per group
    for _ in range(10):
        sch.step()
    assert abs(opt.param_groups[2]['lr'] - 8e-5) < 1e-12


def test_cbcnn_stage1_trains_the_classifier_only():
    t = _bare(examples.CBCNNTrainer, load_config(os.path.join(REPO, 'configs', 'CBCNN_S1.yaml')))
    g = _groups(t)                                                        # Examples/CBCNN.py:13-24
    assert g == [(8192 * 200 + 200, 1.0)]
    assert not any(p.requires_grad for p in t.model.backbone.parameters())


def test_bcnn_stages():
    t1 = _bare(examples.BCNNTrainer, load_config(os.path.join(REPO, 'configs', 'BCNN_S1.yaml')))
    assert _groups(t1) == [(52429000, 1.0)]                               # Examples/BCNN.py:35-36
    t2 = _bare(examples.BCNNTrainer, load_config(os.path.join(REPO, 'configs', 'BCNN_S2.yaml')))
    assert sum(n for n, _ in _groups(t2)) == 67143688                     # :38-39
    assert set(examples.TRAINERS) == {'BCNN', 'CBCNN', 'MPN', 'PeerLearning', 'OSMENet'}


def test_osmenet_param_groups_and_criterion(monkeypatch):
    monkeypatch.setenv('HAWKEYE_ALLOW_RANDOM_INIT', '1')
    cfg = load_config(os.path.join(REPO, 'configs', 'OSMENet.yaml'))
    t = _bare(examples.OSMENetTrainer, cfg)
    g = _groups(t)                                                        # Examples/OSMENet.py:35-42: backbone 0.1 x lr
    assert [m for _, m in g] == [0.1, 1.0]
    n_backbone = sum(p.numel() for p in t.model.backbone.parameters())
    assert g[0][0] == n_backbone and sum(n for n, _ in g) == sum(p.numel() for p in t.model.parameters())
    crit = t.get_criterion(cfg.train.criterion)
    assert crit.lambda_a == 0.5 and crit.use_mamc is True                 # MAMC_loss.py:9-10

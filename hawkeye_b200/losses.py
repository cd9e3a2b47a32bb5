"""Losses of the callers of the hot path that are not plain cross-entropy.

``peer_learning_loss`` follows model/loss/peer_learning_loss.py:5-65 (co-teaching between two networks, Sun et al.,
ICCV 2021): samples on which the two networks DISAGREE are always kept; of the samples on which they agree, each network is
updated on the ``(1 - drop_rate)`` fraction with the smallest loss *under the other network*.  It works on two [N, K] logit
tensors (N = batch size), so it stays in PyTorch: a few microseconds next to a ~25 ms step, and not part of the kernels' path.
"""
import torch
import torch.nn.functional as F


def peer_learning_loss(logits_1, logits_2, labels, drop_rate):
    """-> (loss_1, loss_2), each the mean cross-entropy of its network over the samples it is updated on."""
    pred_1 = logits_1.argmax(dim=1)          # argmax of softmax == argmax of logits (peer_learning_loss.py:15-21)
    pred_2 = logits_2.argmax(dim=1)
    agree = pred_1 == pred_2
    idx_dis = (~agree).nonzero(as_tuple=True)[0]
    idx_agr = agree.nonzero(as_tuple=True)[0]
    sel_1, sel_2 = idx_dis, idx_dis          # sample indices network 1 / 2 is updated on
    if idx_agr.numel() > 0:
        with torch.no_grad():                # ranking only (the reference sorts `.data`, :36-40)
            l1 = F.cross_entropy(logits_1[idx_agr], labels[idx_agr], reduction='none')
            l2 = F.cross_entropy(logits_2[idx_agr], labels[idx_agr], reduction='none')
        keep = int((1 - drop_rate) * idx_agr.numel())                    # :42
        small_1 = idx_agr[torch.argsort(l1)[:keep]]                      # low-loss samples according to network 1
        small_2 = idx_agr[torch.argsort(l2)[:keep]]
        sel_1 = torch.cat((idx_dis, small_2))                            # network 1 learns from network 2's selection (:48-51)
        sel_2 = torch.cat((idx_dis, small_1))
    return F.cross_entropy(logits_1[sel_1], labels[sel_1]), F.cross_entropy(logits_2[sel_2], labels[sel_2])

"""No-op stand-in for tensorboardX (TEST INFRASTRUCTURE ONLY; reference train.py:4)."""


class SummaryWriter:
    def __init__(self, *a, **k):
        pass

    def add_scalar(self, *a, **k):
        pass

    def close(self):
        pass

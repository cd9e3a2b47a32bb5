"""``CfgNode`` used by config.py when yacs (reference config.py:2) is not installed: attribute + item access,
``in``, ``load_cfg``, ``freeze``, ``__str__`` — what the reference's Trainer uses (train.py:41-62,174-180)."""
import yaml


class CfgNode(dict):
    def __init__(self, init_dict=None):
        super().__init__()
        self.__dict__['_frozen'] = False
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.__dict__.get('_frozen'):
            raise AttributeError('CfgNode is frozen')
        self[name] = value

    @classmethod
    def load_cfg(cls, f):
        return cls(yaml.safe_load(f.read() if hasattr(f, 'read') else f))

    def freeze(self):
        self.__dict__['_frozen'] = True
        for v in self.values():
            if isinstance(v, CfgNode):
                v.freeze()

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, CfgNode) else v) for k, v in self.items()}

    def __str__(self):
        return yaml.safe_dump(self.to_dict(), default_flow_style=False)

    __repr__ = __str__

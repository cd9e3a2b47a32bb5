"""Minimal stand-in for ``yacs.config.CfgNode`` (TEST INFRASTRUCTURE ONLY).

The reference imports yacs in ``config.py:2`` and ``utils/utils.py:7``; yacs is
not installed in this image and there is no network.  This shim is put on
``sys.path`` by ``oracle/ref_harness.py`` only when the real package fails to
import, so that the unmodified reference tree can be imported as the oracle.
It supports what the reference uses: attribute + item access, ``in``,
``load_cfg(file)``, ``freeze()``, ``__str__``.
"""
import yaml


class CfgNode(dict):
    def __init__(self, init_dict=None):
        super().__init__()
        self.__dict__["_frozen"] = False
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.__dict__.get("_frozen"):
            raise AttributeError("CfgNode is frozen")
        self[name] = value

    @classmethod
    def load_cfg(cls, f):
        if hasattr(f, "read"):
            f = f.read()
        return cls(yaml.safe_load(f))

    def freeze(self):
        self.__dict__["_frozen"] = True
        for v in self.values():
            if isinstance(v, CfgNode):
                v.freeze()

    def defrost(self):
        self.__dict__["_frozen"] = False
        for v in self.values():
            if isinstance(v, CfgNode):
                v.defrost()

    def _as_dict(self):
        return {k: (v._as_dict() if isinstance(v, CfgNode) else v) for k, v in self.items()}

    def __str__(self):
        return yaml.safe_dump(self._as_dict(), default_flow_style=False)

    __repr__ = __str__

"""Minimal stand-in for mmcv.Config.fromfile (mmcv 0.2.14 is not vendored by the reference).

Executes a python config file and exposes its top-level names as an attribute dict whose nested
dicts are attribute dicts too.  `hasattr(cfg.test_cfg, 'flownet2')`-style feature switches
(panoptic_fusetrack.py:59-64,90-91,513,560) work because missing keys raise AttributeError.
"""
import os


class ConfigDict(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError("'ConfigDict' object has no attribute '%s'" % name)

    def __setattr__(self, name, value):
        self[name] = value


def _wrap(v):
    if isinstance(v, dict):
        return ConfigDict({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, (list, tuple)):
        return type(v)(_wrap(x) for x in v)
    return v


class Config(object):
    def __init__(self, cfg_dict, filename=None, text=""):
        object.__setattr__(self, "_cfg_dict", _wrap(cfg_dict))
        object.__setattr__(self, "filename", filename)
        object.__setattr__(self, "text", text)

    @staticmethod
    def fromfile(filename):
        filename = os.path.abspath(os.path.expanduser(filename))
        with open(filename) as f:
            text = f.read()
        ns = {}
        exec(compile(text, filename, "exec"), ns)
        cfg = {k: v for k, v in ns.items() if not k.startswith("__") and not callable(v) and not hasattr(v, "__loader__")}
        return Config(cfg, filename, text)

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def get(self, name, default=None):
        return self._cfg_dict.get(name, default)

    def __contains__(self, name):
        return name in self._cfg_dict

    def __repr__(self):
        return "Config (path: %s): %r" % (self.filename, self._cfg_dict)

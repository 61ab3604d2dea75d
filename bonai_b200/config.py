"""Python-file configs with `_base_` inheritance, `_delete_` keys and attribute + item access:
the config surface of the reference (mmcv.Config.fromfile as used by tools/train.py:71-73), so
configs/loft_foa/*.py and their four `_base_` files load unchanged.  Configs are *executed*
(they contain expressions and loops, configs/_base_/datasets/bonai_instance.py:33-38)."""
import ast
import copy
import os.path as osp
import types
from argparse import Action


class ConfigDict(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(f"'ConfigDict' object has no attribute '{name}'")

    def __setattr__(self, name, value):
        self[name] = value

    def __deepcopy__(self, memo):
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def _wrap(v):
    if isinstance(v, dict):
        return ConfigDict({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, list):
        return [_wrap(x) for x in v]
    if isinstance(v, tuple):
        return tuple(_wrap(x) for x in v)
    return v


def _merge(child, base):
    out = dict(base)
    for k, v in child.items():
        if isinstance(v, dict) and not v.get('_delete_', False) and isinstance(out.get(k), dict):
            out[k] = _merge(v, out[k])
        else:
            if isinstance(v, dict):
                v = {a: b for a, b in v.items() if a != '_delete_'}
            out[k] = v
    return out


def _load(filename):
    filename = osp.abspath(osp.expanduser(filename))
    if not osp.isfile(filename):
        raise FileNotFoundError(f'config file {filename} does not exist')
    if not filename.endswith('.py'):
        raise IOError('Only .py configs are supported')
    with open(filename) as f:
        src = f.read()
    ns = {'__file__': filename}
    exec(compile(src, filename, 'exec'), ns)
    cfg = {k: v for k, v in ns.items()
           if not k.startswith('__')
           and not isinstance(v, (types.ModuleType, types.FunctionType, type))}
    bases = cfg.pop('_base_', None)
    if bases is not None:
        bases = bases if isinstance(bases, list) else [bases]
        merged = {}
        for b in bases:
            bd = _load(osp.join(osp.dirname(filename), b))
            dup = merged.keys() & bd.keys()
            if dup:
                raise KeyError(f'Duplicate key is not allowed among bases: {sorted(dup)}')
            merged.update(bd)
        cfg = _merge(cfg, merged)
    return cfg


class Config:
    def __init__(self, cfg_dict=None, filename=None):
        if cfg_dict is not None and not isinstance(cfg_dict, dict):
            raise TypeError(f'cfg_dict must be a dict, but got {type(cfg_dict)}')
        object.__setattr__(self, '_cfg_dict', _wrap(cfg_dict or {}))
        object.__setattr__(self, '_filename', filename)

    @staticmethod
    def fromfile(filename, use_predefined_variables=True):
        return Config(_load(filename), filename=filename)

    @property
    def filename(self):
        return self._filename

    def __repr__(self):
        return f'Config (path: {self._filename}): {dict.__repr__(self._cfg_dict)}'

    def __len__(self):
        return len(self._cfg_dict)

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __setattr__(self, name, value):
        self._cfg_dict[name] = _wrap(value)

    def __setitem__(self, name, value):
        self._cfg_dict[name] = _wrap(value)

    def __contains__(self, name):
        return name in self._cfg_dict

    def __iter__(self):
        return iter(self._cfg_dict)

    def get(self, key, default=None):
        return self._cfg_dict.get(key, default)

    def merge_from_dict(self, options):
        nested = {}
        for full_key, v in options.items():
            cur = nested
            keys = full_key.split('.')
            for k in keys[:-1]:
                cur = cur.setdefault(k, {})
            cur[keys[-1]] = v
        object.__setattr__(self, '_cfg_dict', _wrap(_merge(nested, dict(self._cfg_dict))))


class DictAction(Action):
    """argparse action for `--options k=v ...` (tools/train.py:49-50)."""

    def __call__(self, parser, namespace, values, option_string=None):
        opts = {}
        for kv in values:
            k, v = kv.split('=', maxsplit=1)
            try:
                v = ast.literal_eval(v)
            except (ValueError, SyntaxError):
                pass
            opts[k] = v
        setattr(namespace, self.dest, opts)

import ast
import copy
import os.path as osp
import types
from argparse import Action


class ConfigDict(dict):
    """dict with attribute access (stand-in for addict.Dict as used by mmcv.Config)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def __deepcopy__(self, memo):
        return ConfigDict({k: copy.deepcopy(v, memo) for k, v in self.items()})


def _to_cfgdict(v):
    if isinstance(v, dict):
        return ConfigDict({k: _to_cfgdict(x) for k, x in v.items()})
    if isinstance(v, list):
        return [_to_cfgdict(x) for x in v]
    if isinstance(v, tuple):
        return tuple(_to_cfgdict(x) for x in v)
    return v


def _merge_a_into_b(a, b):
    b = b.copy()
    for k, v in a.items():
        if isinstance(v, dict) and k in b and not v.get('_delete_', False):
            if not isinstance(b[k], dict):
                raise TypeError(f'{k}: cannot merge dict into {type(b[k])}')
            b[k] = _merge_a_into_b(v, b[k])
        else:
            if isinstance(v, dict):
                v = {kk: vv for kk, vv in v.items() if kk != '_delete_'}
            b[k] = v
    return b


def _file2dict(filename):
    filename = osp.abspath(filename)
    with open(filename) as f:
        src = f.read()
    ns = {}
    exec(compile(src, filename, 'exec'), ns)
    cfg = {k: v for k, v in ns.items()
           if not k.startswith('__') and not isinstance(v, (types.ModuleType, types.FunctionType))
           and not isinstance(v, type)}
    if '_base_' in cfg:
        base = cfg.pop('_base_')
        base = base if isinstance(base, list) else [base]
        merged = {}
        for b in base:
            bd = _file2dict(osp.join(osp.dirname(filename), b))
            dup = merged.keys() & bd.keys()
            if dup:
                raise KeyError(f'Duplicate key in base configs: {dup}')
            merged.update(bd)
        cfg = _merge_a_into_b(cfg, merged)
    return cfg


class Config:
    def __init__(self, cfg_dict=None, filename=None):
        object.__setattr__(self, '_cfg_dict', _to_cfgdict(cfg_dict or {}))
        object.__setattr__(self, '_filename', filename)

    @staticmethod
    def fromfile(filename, use_predefined_variables=True):
        return Config(_file2dict(filename), filename=filename)

    @property
    def filename(self):
        return self._filename

    @property
    def pretty_text(self):
        return repr(self._cfg_dict)

    def __getattr__(self, name):
        return getattr(self._cfg_dict, name)

    def __getitem__(self, name):
        return self._cfg_dict[name]

    def __setattr__(self, name, value):
        self._cfg_dict[name] = _to_cfgdict(value)

    def __setitem__(self, name, value):
        self._cfg_dict[name] = _to_cfgdict(value)

    def __contains__(self, name):
        return name in self._cfg_dict

    def __iter__(self):
        return iter(self._cfg_dict)

    def get(self, key, default=None):
        return self._cfg_dict.get(key, default)

    def merge_from_dict(self, options):
        d = {}
        for full_key, v in options.items():
            cur = d
            keys = full_key.split('.')
            for k in keys[:-1]:
                cur = cur.setdefault(k, {})
            cur[keys[-1]] = v
        object.__setattr__(self, '_cfg_dict',
                           _to_cfgdict(_merge_a_into_b(d, dict(self._cfg_dict))))


class DictAction(Action):
    def __call__(self, parser, namespace, values, option_string=None):
        opts = {}
        for kv in values:
            k, v = kv.split('=', 1)
            try:
                v = ast.literal_eval(v)
            except Exception:
                pass
            opts[k] = v
        setattr(namespace, self.dest, opts)

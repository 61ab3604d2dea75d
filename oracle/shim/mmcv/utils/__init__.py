import inspect
import logging

from _shim_dummy import install_getattr as _ig


def is_str(x):
    return isinstance(x, str)


def is_seq_of(seq, expected_type, seq_type=None):
    import collections.abc as abc
    exp = seq_type or abc.Sequence
    if not isinstance(seq, exp):
        return False
    return all(isinstance(i, expected_type) for i in seq)


def is_list_of(seq, t):
    return is_seq_of(seq, t, list)


def is_tuple_of(seq, t):
    return is_seq_of(seq, t, tuple)


class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    def __len__(self):
        return len(self._module_dict)

    def __contains__(self, key):
        return self.get(key) is not None

    @property
    def name(self):
        return self._name

    @property
    def module_dict(self):
        return self._module_dict

    def get(self, key):
        return self._module_dict.get(key, None)

    def _register_module(self, module_class, module_name=None, force=False):
        if not inspect.isclass(module_class):
            raise TypeError(f'module must be a class, but got {type(module_class)}')
        if module_name is None:
            module_name = module_class.__name__
        if not force and module_name in self._module_dict:
            raise KeyError(f'{module_name} is already registered in {self.name}')
        self._module_dict[module_name] = module_class

    def register_module(self, name=None, force=False, module=None):
        # old-style bare decorator: @X.register_module
        if inspect.isclass(name):
            self._register_module(name, force=force)
            return name
        if module is not None:
            self._register_module(module, module_name=name, force=force)
            return module

        def _register(cls):
            self._register_module(cls, module_name=name, force=force)
            return cls

        return _register


def build_from_cfg(cfg, registry, default_args=None):
    if not isinstance(cfg, dict):
        raise TypeError(f'cfg must be a dict, but got {type(cfg)}')
    if 'type' not in cfg:
        raise KeyError(f'`cfg` must contain the key "type", but got {cfg}')
    args = dict(cfg)
    obj_type = args.pop('type')
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError(f'{obj_type} is not in the {registry.name} registry')
    elif inspect.isclass(obj_type):
        obj_cls = obj_type
    else:
        raise TypeError(f'type must be a str or valid type, but got {type(obj_type)}')
    if default_args is not None:
        for name, value in default_args.items():
            args.setdefault(name, value)
    return obj_cls(**args)


def get_logger(name, log_file=None, log_level=logging.INFO):
    return logging.getLogger(name)


def print_log(msg, logger=None, level=logging.INFO):
    if logger == 'silent':
        return
    if logger is None:
        print(msg)
    elif isinstance(logger, logging.Logger):
        logger.log(level, msg)


def get_build_config():
    return ''


_ig(globals(), 'mmcv.utils')

"""Permissive placeholders for names the reference imports but the LOFT path never executes."""
import types


class _DummyMeta(type):
    def __getattr__(cls, name):
        if name.startswith('__'):
            raise AttributeError(name)
        return make_dummy(f'{cls.__name__}.{name}')


def make_dummy(name):
    def _init(self, *a, **k):
        raise NotImplementedError(f'oracle shim: {name} is a placeholder, not implemented')

    return _DummyMeta(name.split('.')[-1], (), {'__init__': _init, '_shim_dummy': True})


def install_getattr(module_globals, modname):
    def __getattr__(name):
        if name.startswith('__'):
            raise AttributeError(name)
        d = make_dummy(f'{modname}.{name}')
        module_globals[name] = d
        return d

    module_globals['__getattr__'] = __getattr__

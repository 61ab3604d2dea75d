from _shim_dummy import install_getattr as _ig

_ig(globals(), 'pycocotools')

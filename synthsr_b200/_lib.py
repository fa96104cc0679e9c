"""ctypes binding of libsynthsr_b200.so.

The prototypes are parsed from include/synthsr_b200.h, so the Python side can never drift from the C ABI.  There is
no fallback: if the library is missing or a call fails, an exception is raised (the product path must fail loudly).
"""
import ctypes
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), 'include', 'synthsr_b200.h')
LIB_PATH = os.path.join(_HERE, 'lib', 'libsynthsr_b200.so')

_SCALARS = {
    'int': ctypes.c_int,
    'float': ctypes.c_float,
    'double': ctypes.c_double,
    'long long': ctypes.c_longlong,
    'unsigned long long': ctypes.c_ulonglong,
    'unsigned int': ctypes.c_uint,
}


class SsrError(RuntimeError):
    pass


def parse_header(path=HEADER):
    """-> {name: (restype_str, [(ctype_str, argname), ...])} for every prototype in the header."""
    text = open(path).read()
    text = re.sub(r'/\*.*?\*/', ' ', text, flags=re.S)
    text = re.sub(r'//[^\n]*', ' ', text)
    text = re.sub(r'^\s*#.*$', ' ', text, flags=re.M)
    text = text.replace('extern "C" {', ' ')
    protos = {}
    for m in re.finditer(r'([A-Za-z_][A-Za-z0-9_ \*]*?)\s*\b(ssr_[a-z0-9_]+)\s*\(([^)]*)\)\s*;', text):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        arglist = []
        if args and args != 'void':
            for a in args.split(','):
                a = ' '.join(a.split())
                mm = re.match(r'(.*?)([A-Za-z_][A-Za-z0-9_]*)$', a)
                arglist.append((mm.group(1).strip(), mm.group(2)))
        protos[name] = (ret, arglist)
    return protos


def _ctype(tstr):
    t = tstr.replace('const ', '').strip()
    if t.endswith('*'):
        return ctypes.c_char_p if t == 'char*' else ctypes.c_void_p
    return _SCALARS[t]


class _Lib:
    def __init__(self):
        self._dll = None
        self.protos = parse_header()

    def _load(self):
        if self._dll is not None:
            return self._dll
        if not os.path.exists(LIB_PATH):
            raise SsrError('libsynthsr_b200.so not found at %s -- run `python -m synthsr_b200.build` '
                           '(there is no CPU fallback)' % LIB_PATH)
        try:
            import torch  # noqa: F401  (loads the CUDA runtime the library shares with torch)
        except Exception:
            pass
        dll = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
        for name, (ret, args) in self.protos.items():
            fn = getattr(dll, name)   # AttributeError if the .so does not export a declared symbol
            fn.restype = _ctype(ret) if ret != 'void' else None
            fn.argtypes = [_ctype(t) for t, _ in args]
        self._dll = dll
        return dll

    def raw(self, name):
        return getattr(self._load(), name)

    def __getattr__(self, name):
        if name.startswith('_') or name == 'protos':
            raise AttributeError(name)
        dll = self._load()
        fn = getattr(dll, name)
        ret = self.protos[name][0]

        def call(*args):
            conv = []
            for a in args:
                if hasattr(a, 'data_ptr'):
                    conv.append(a.data_ptr())
                else:
                    conv.append(a)
            r = fn(*conv)
            if ret == 'int' and r < 0:
                raise SsrError('%s failed (%d): %s' % (name, r, dll.ssr_last_error().decode()))
            return r

        call.__name__ = name
        setattr(self, name, call)
        return call


lib = _Lib()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream

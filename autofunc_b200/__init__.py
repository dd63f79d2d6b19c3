"""autofunc_b200 -- B200-native dense-network differentiation path of unixpickle/autofunc.

`autofunc_b200.api` mirrors the reference's Func/RFunc/Batcher/RBatcher + Result/RResult interface;
`autofunc_b200._lib` is the ctypes binding of the C ABI in include/autofunc_b200.h.
"""
from . import _lib  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import Context, DeviceVec, MLPPlan, context, set_context  # noqa: F401
from .slices import concat, concat_r, slice_, slice_r, repeat, repeat_r, concat_batched, concat_batched_r  # noqa: F401,E402
from . import api as _api  # noqa: E402
for _n in ("concat", "concat_r", "slice_", "slice_r", "repeat", "repeat_r", "concat_batched", "concat_batched_r"):
    setattr(_api, _n, globals()[_n])  # the operator namespace tests and callers use is `api`
from . import funcs as _funcs  # noqa: E402
from .funcs import *  # noqa: F401,F403,E402
for _n in _funcs.__all__:
    setattr(_api, _n, getattr(_funcs, _n))

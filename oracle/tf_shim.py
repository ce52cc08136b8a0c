"""A minimal ``tensorflow`` namespace over torch, so that the reference's TensorFlow ORIGINAL of the
projection (``/root/reference/dpc/util/point_cloud.py``, the only complete statement of the
point-feature / RGB branch) can be EXECUTED UNMODIFIED in the build container, where TensorFlow
is not installed.

TEST INFRASTRUCTURE (see oracle/__init__.py): used by ``oracle.ref_loader.ref_project_rgb_tf`` to
pin ``oracle/rgb.py`` and to make ``tests/golden/rgb_tf.npz``.  Never imported by the product and
never on the GPU box (the reference tree is absent there).

Only the TF-1 graph ops that file calls on the ``pointcloud_project_fast`` path are provided, each
with the documented TF semantics (op name -> what it does here):

  expand_dims, reshape, tile, concat, slice, reverse, cast, floor, range, constant, shape,
  logical_and, reduce_all, reduce_max, reduce_sum, clip_by_value, add_n, stop_gradient
      the elementwise / shape op of the same name; ``slice`` returns a fresh tensor
  boolean_mask(t, m)           rows of ``t`` where the 1-D mask ``m`` holds
  scatter_nd(idx, upd, shape)  zeros(shape) with ``upd`` ADDED at ``idx`` (duplicates accumulate;
                               ``idx`` may address a prefix of the dimensions)
  nn.conv3d(x, f, strides, "SAME")  x NDHWC, f [kd,kh,kw,cin,cout], unit strides, odd taps:
                               cross-correlation with k//2 zero padding each side

TF tensors are immutable: ``a += b`` rebinds ``a`` to a new tensor.  torch's ``+=`` writes in
place (and would corrupt values autograd still needs: the reference does ``xs /= zs`` and then
``zs -= camera_distance``), so every tensor that enters or leaves this namespace is a ``TFTensor``
whose augmented assignments are out of place.

The shim itself is validated, not trusted: ``tests/test_rgb.py::test_tf_shim_reproduces_the_torch_port``
runs the TF file through it WITHOUT features and requires the reference's own torch port
(``ref_loader.ref_project``) bit for bit on every output and gradient -- which exercises every op
above except the feature-only lines (``expand_dims(w) * rgb``, the 5-D ``scatter_nd``, the division).
"""
import builtins
import types

import torch
import torch.nn.functional as F


class TFTensor(torch.Tensor):
    """torch.Tensor with TensorFlow's value semantics for ``+= -= *= /=``."""

    def __iadd__(self, other):
        return self + other

    def __isub__(self, other):
        return self - other

    def __imul__(self, other):
        return self * other

    def __itruediv__(self, other):
        return self / other


def wrap(t):
    return None if t is None else t.as_subclass(TFTensor)


def unwrap(t):
    return None if t is None else t.as_subclass(torch.Tensor)


float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
int64 = torch.int64


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x)


def constant(value, dtype=None):
    return wrap(torch.tensor(value, dtype=dtype))


def convert_to_tensor(value, dtype=None):
    return wrap(torch.as_tensor(value, dtype=dtype))


def shape(x):
    return tuple(x.shape)


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def reshape(x, shp):
    return x.reshape([int(s) for s in shp])


def tile(x, multiples):
    return x.repeat(*[int(m) for m in multiples])


def concat(values, axis):
    return torch.cat(list(values), dim=axis)


def slice(x, begin, size):                                            # noqa: A001 (TF's name)
    idx = tuple(builtins.slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))
    return x[idx].clone()


def reverse(x, axis):
    return torch.flip(x, [int(a) for a in axis])


def cast(x, dtype):
    return x.to(dtype)


def floor(x):
    return torch.floor(x)


def range(start, limit, delta=1):                                     # noqa: A001 (TF's name)
    return wrap(torch.arange(int(start), int(limit), int(delta), dtype=torch.int32))


def logical_and(a, b):
    return torch.logical_and(a, b)


def reduce_all(x, axis):
    return torch.all(x, dim=axis)


def reduce_max(x, axis):
    return torch.amax(x, dim=[int(a) for a in axis])


def reduce_sum(x, axis=None):
    return x.sum() if axis is None else x.sum(dim=axis)


def clip_by_value(x, lo, hi):
    return torch.clamp(x, lo, hi)


def add_n(values):
    out = values[0]
    for v in values[1:]:
        out = out + v
    return out


def stop_gradient(x):
    return x.detach()


def boolean_mask(tensor, mask):
    assert mask.dim() == 1 and mask.dtype == torch.bool
    return tensor[mask]


def scatter_nd(indices, updates, shape):                              # noqa: A002 (TF's name)
    shp = [int(s) for s in shape]
    idx = indices.long()
    assert idx.dim() == 2 and idx.shape[1] <= len(shp)
    out = torch.zeros(shp, dtype=updates.dtype).as_subclass(TFTensor)
    return out.index_put(tuple(idx.unbind(1)), updates, accumulate=True)


def _conv3d(input, filter, strides, padding):                         # noqa: A002 (TF's names)
    assert list(strides) == [1, 1, 1, 1, 1] and padding == "SAME"
    kd, kh, kw = (int(k) for k in filter.shape[:3])
    assert kd % 2 == kh % 2 == kw % 2 == 1, "SAME padding is symmetric only for odd taps"
    w = filter.to(input.dtype).permute(4, 3, 0, 1, 2)                 # -> [cout, cin, kd, kh, kw]
    y = F.conv3d(input.permute(0, 4, 1, 2, 3), w, stride=1, padding=(kd // 2, kh // 2, kw // 2))
    return y.permute(0, 2, 3, 4, 1)


nn = types.SimpleNamespace(conv3d=_conv3d)

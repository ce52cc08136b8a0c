"""Host-side Gaussian taps: mirror of the reference's util/gauss_kernel.py.

The taps are an *input* of the projection path (the model rebuilds them every
step from the sigma schedule, model_pc_to.py:171-179), so they are computed on
the host exactly as the reference does -- fp32 torch ops on CPU tensors -- and
handed to the CUDA kernels as constant-bank arguments.
"""
import torch


def gauss_kernel_1d(l, sig):
    """Normalised 1-D Gaussian of length ``l`` (util/gauss_kernel.py:5-11)."""
    support = torch.arange(-l // 2 + 1., l // 2 + 1)
    weights = torch.exp(-support ** 2 / (2. * sig ** 2))
    return weights / weights.sum()


def separable_kernels(kernel):
    """1-D taps -> the [X, Y, Z] conv3d kernel list (util/gauss_kernel.py:27-32)."""
    n = kernel.shape[0]
    return [kernel.reshape((1, 1, 1, 1, n)), kernel.reshape((1, 1, 1, n, 1)),
            kernel.reshape((1, 1, n, 1, 1))]


def smoothing_kernel(cfg, sigma):
    """The kernel list ``pointcloud_project_fast`` takes (util/gauss_kernel.py:35-55).

    With ``vox_size_z != -1`` the Z taps are shorter and narrower by the ratio
    vox_size_z / vox_size (:38-47).  The reference then reshapes them with the
    XY length (:49) and raises; the intended length is used here.
    Only the separable filter exists (``pc_separable_gauss_filter: true`` in
    every config; the reference's non-separable branch returns an unbound name).
    """
    size = cfg.pc_gauss_kernel_size
    taps = gauss_kernel_1d(size, sigma)
    vox_size_z = getattr(cfg, "vox_size_z", -1)
    if vox_size_z != -1:
        ratio = vox_size_z / cfg.vox_size
        size_z = int((size * ratio) // 1)
        if size_z % 2 == 0:
            size_z += 1
        taps_z = gauss_kernel_1d(size_z, sigma * ratio)
        return [taps.reshape((1, 1, 1, 1, size)), taps.reshape((1, 1, 1, size, 1)),
                taps_z.reshape((1, 1, size_z, 1, 1))]
    if not getattr(cfg, "pc_separable_gauss_filter", True):
        raise NotImplementedError("only the separable Gaussian filter is supported "
                                  "(the reference's non-separable branch is unreachable)")
    return separable_kernels(taps)

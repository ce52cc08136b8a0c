"""CPU oracle for the differentiable point-cloud projection path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline -- never as the thing shipped.  The product path
(``pytorch_unsup_pc_b200``) never imports this package and fails loudly when
its CUDA library is missing.

Contents
--------
``closed_form``  torch-on-CPU restatement of the reference algorithm
                 (reference dtype flow: fp32 quaternion normalise + first
                 Hamilton product, fp64 afterwards), each function citing the
                 reference file:line it follows.  Gradients come from torch
                 autograd over the restatement, exactly as the reference gets
                 its own gradients.
``ref_loader``   imports the *real* reference from ``/root/reference`` (only
                 present in the build container, never on the GPU box) so the
                 restatement can be pinned against it and golden vectors can
                 be generated (``tests/golden/make_golden.py``).
``config``       the handful of config keys the path reads, with the
                 reference's defaults.
``loss`` ``replicas`` ``render_loss`` ``chamfer``
                 restatements of the next rows (candidate-selection
                 loss, tf_repeat_0 + point dropout, their composition with the
                 projection, the Chamfer nearest neighbour), each pinned against the
                 reference's own functions executed live (loss, argmin, indices and
                 forward values bit for bit; Chamfer distances to 1 ulp of torch's
                 CPU sqrt; seeded sweeps in tests/) and by reference-made golden
                 vectors.
``rgb``          restatement of the point-feature branch from the reference's
                 TensorFlow original (its torch port does not run), pinned against
                 that TF source executed unmodified through ``tf_shim``.
``tf_shim``      a minimal ``tensorflow`` namespace over torch (the ~25 TF-1 ops
                 util/point_cloud.py calls), validated bit for bit against the
                 reference's torch port on the occupancy path.

Parity status: the reference ships no golden vectors;
the oracle is pinned by executing the reference itself in the build container
(``tests/test_oracle_pinning.py`` live, ``tests/golden/*.npz`` committed).
"""

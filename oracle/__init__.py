"""CPU oracle for the DeepPrior++ hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (NumPy fp64/fp32 + torch-CPU fp32 + cv2), the
algorithms of the reference (moberweger/deep-prior-pp) for the single hot path this
repository re-implements in CUDA: depth-crop augmentation -> ResNet/PoseRegNet forward,
loss, backward, ADAM.  Every function cites the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  The product package never does.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures for this path
(SURVEY.md section 4), and its own implementation (Theano 0.9 / Python 2.7) cannot be
executed in this environment.  The pins are (a) cv2 4.13.0 as installed here for the warp
index rules (``oracle/augment.py`` checks its NumPy index model against cv2 itself) and
(b) the committed vectors under ``tests/golden/`` produced by ``tests/golden/make_golden.py``.
"""

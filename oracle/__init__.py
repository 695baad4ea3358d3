"""CPU oracle for the DeepPrior++ hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (NumPy fp64/fp32 + torch-CPU fp32 + cv2), the
algorithms of the reference (moberweger/deep-prior-pp) for the single hot path this
repository re-implements in CUDA: depth-crop augmentation -> ResNet/PoseRegNet forward,
loss, backward, ADAM; the inference cascade (CoM refinement -> re-crop -> pose); pose sampling.
Every function cites the reference file:line it follows.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it (``__graft_entry__.build()`` additionally checks that
``ref_harness`` can load the reference sources where they exist).  The product package never does.

PARITY STATUS
* Augmentation, crop geometry, projections, cascade crops, sampleRandomPoses (oracle/augment.py, oracle/cascade.py):
  PINNED AGAINST THE REFERENCE'S OWN CODE.  oracle/ref_harness.py executes the reference's Python sources from
  /root/reference (py2 -> py3 pass in memory: print statements, classic integer division, xrange; nothing copied) in
  the build container; tests/golden/make_reference_vectors.py stored its outputs in tests/golden/reference_pins.npz and
  tests/test_reference_pins.py compares the oracle with them (and, where /root/reference exists, live on fresh
  seeds): every index / integer result bit-exact, float results within a few float32 ulps (the fixture ran under
  NumPy 2, the oracle restates the reference-era NumPy 1.x promotion rules).  The pixel rules of the
  warps are additionally pinned to cv2 4.13.0 itself (tests/test_oracle_warp.py, tests/test_oracle_cascade.py).
* Network STRUCTURE and INITIALISATION (oracle/nets.py builders and the product's net classes): PINNED against the
  reference's own constructors - oracle/ref_harness.py::describe_reference_net runs net/resnet.py, net/poseregnet.py,
  net/scalenet.py and every layer class with an inert stand-in for theano (the numeric work - wiring, layer numbers,
  dimension arithmetic, weight initialisation, order of random draws, parameter order - runs for real);
  tests/golden/reference_nets.json holds the result, tests/test_reference_pins.py checks layer lists, dimensions,
  parameter names / order and the sha1 of every initial weight tensor: bit-identical.
* Network ARITHMETIC - forward ops, cost, gradients, ADAM (oracle/nets.py): pinned against the reference's own
  Python code EVALUATED EAGERLY (oracle/eager_theano.py: a stand-in for Theano whose calls compute, torch-CPU float64
  + autograd; oracle/ref_harness.py::run_reference_net runs net/*.py, PoseRegNetTrainer.setupFunctions and
  Optimizer.ADAM through it) -> tests/golden/reference_net_eval.npz, tests/test_reference_pins.py: outputs, every
  layer's output, cost, all gradients, BatchNorm EMA, one ADAM step.  Theano's PRIMITIVES (conv2d, pool_2d, var, ...)
  are NOT pinned: Theano 0.9 cannot be installed or run here (Python 3.12, no network) and the reference ships no
  tests or fixtures; stand-in and oracle both follow the documented semantics.
"""

"""The two Theano primitives whose semantics the oracle (and oracle/eager_theano.py) take from Theano 0.9's
documentation rather than from executable reference code - stated independently with SciPy / NumPy:

 * ``theano.tensor.nnet.conv2d`` is a TRUE convolution (``filter_flip=True``): out[o] = sum_c x[c] (*) W[o, c], which is
   ``scipy.signal.convolve2d``; ``border_mode='valid'`` keeps full overlaps, ``'half'`` pads k // 2 (= SciPy's 'same'
   for odd kernels), ``subsample=(s, s)`` keeps every s-th output of the stride-1 result;
 * ``pool_2d(ds, ignore_border=True, mode='max')`` takes the maximum of non-overlapping ds x ds blocks and drops the
   incomplete border blocks.
The oracle's layer forward (oracle/nets.py) must agree with these statements."""
import numpy as np
import pytest
import torch

scipy_signal = pytest.importorskip('scipy.signal')

from oracle import nets as ON  # noqa: E402


def _oracle_conv(x, W, b, stride, border, pool=1):
    net = ON.OracleNet(np.random.RandomState(0), x.shape)
    cin, hw, nf, k = x.shape[1], x.shape[2:], W.shape[0], W.shape[2]
    if pool > 1 or border == 'valid':
        v, _, _ = net.add_convpool(0, cin, hw, nf, k, pool, border, 'None', 'He')
    else:
        v, _, _ = net.add_conv(0, cin, hw, nf, k, stride, border)
    net.layers[-1].params = [torch.from_numpy(W), torch.from_numpy(b)]
    net.out_vid = v
    with torch.no_grad():
        o, _ = net.forward(torch.from_numpy(x), deterministic=True)
    return o.numpy()


@pytest.mark.parametrize('k,border,stride', [(5, 'valid', 1), (3, 'half', 1), (1, 'half', 2), (5, 'half', 1), (3, 'valid', 1)])
def test_conv2d_is_a_true_convolution(k, border, stride):
    rng = np.random.RandomState(k * 7 + stride)
    x = rng.randn(2, 3, 12, 14)
    W = rng.randn(4, 3, k, k)
    b = rng.randn(4)
    got = _oracle_conv(x, W, b, stride, border)
    mode = 'valid' if border == 'valid' else 'same'
    want = np.stack([np.stack([sum(scipy_signal.convolve2d(x[n, c], W[o, c], mode=mode) for c in range(3)) + b[o]
                               for o in range(4)]) for n in range(2)])
    want = want[:, :, ::stride, ::stride]
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize('H,pool', [(13, 2), (12, 4), (11, 3), (9, 1)])
def test_pool_2d_ignores_incomplete_border_blocks(H, pool):
    rng = np.random.RandomState(H)
    x = rng.randn(2, 3, H + 2, H + 4)
    W = np.zeros((3, 3, 3, 3))
    for c in range(3):
        W[c, c, 1, 1] = 1.0                     # identity convolution ('valid' trims one pixel per side)
    b = rng.randn(3)
    got = _oracle_conv(x, W, b, 1, 'valid', pool=pool)
    core = x[:, :, 1:-1, 1:-1]
    hp, wp = core.shape[2] // pool, core.shape[3] // pool
    blocks = core[:, :, :hp * pool, :wp * pool].reshape(2, 3, hp, pool, wp, pool)
    want = blocks.max(axis=(3, 5)) + b.reshape(1, 3, 1, 1)      # the bias is added AFTER pooling (convpoollayer.py:276)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)

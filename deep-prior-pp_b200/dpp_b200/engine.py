"""Graph executor: lowers a network object built by the reference-surface classes (``net.*``) to
calls into libdpp_b200.so and owns the device buffers (torch CUDA tensors used as storage only).

What it replaces in the reference: the three ``theano.function`` compilations -
``compute_output`` (src/net/netbase.py:257-282), ``train_model`` and the validation functions
(src/trainer/poseregnettrainer.py:146-209) - i.e. Theano's graph compilation, ``T.grad`` and the
``updates`` mechanism.  The graph recorded by ``net.sym.Sym`` is pattern-matched onto the fused
kernels:

  BN -> ReLU -> ConvLayer          => dpp_conv2d_fwd with in_bn prologue   (BN/ReLU never stored)
  ConvLayer -> (+ identity/shortcut) => residual epilogue of dpp_conv2d_fwd
  tensor -> BN                      => fp64 sum/sumsq in the producer's epilogue (out_stats)
  BN -> ReLU -> flatten -> Hidden   => dpp_bn_apply (materialised once), dpp_fc_fwd
  Hidden -> Dropout                 => mask / 0.7 scale in dpp_fc_fwd's epilogue

The backward program is the reverse walk of the forward op list (dgrad/wgrad/bn_bwd_apply), i.e.
what ``T.grad(cost, params)`` (poseregnettrainer.py:111) produced symbolically.

There is NO CPU path: constructing an Engine without CUDA raises.
"""
import ctypes as C
import os
import numpy as np

from .lib import lib, BnRef, ConvDesc, BnEmaItem, PackItem, WgradLayer, DppError, DPP_ENOTSUP


def _torch():
    import torch
    return torch


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class Tensor(object):
    """A materialised activation: NHWC (4-D) or (B, n) (2-D) fp32 device buffer."""

    def __init__(self, name, shape_nchw, batch=None):
        self.name = name
        self.shape = tuple(int(v) for v in shape_nchw)    # reference (NCHW or (B,n)) shape
        if batch is not None:
            self.shape = (int(batch),) + self.shape[1:]   # this device's share of the minibatch
        self.buf = None
        self.grad = None
        self.bn = None              # BatchNormLayer normalising this tensor (if any)
        self.is_input = False
        self.chw = None             # (C,H,W) if a 2-D tensor is the flatten of a 4-D one

    @property
    def numel(self):
        return int(np.prod(self.shape))

    @property
    def pixels(self):
        s = self.shape
        return s[0] * s[2] * s[3]


class Slot(object):
    def __init__(self, var, arena, offset):
        self.var = var
        self.arena = arena          # 'w' (trainable) or 'r' (non-trained)
        self.offset = offset
        self.shape = tuple(var.shape)
        self.size = int(np.prod(self.shape)) if len(self.shape) else 1
        self.chw = None             # FC rows permuted from (c,h,w) to (h,w,c)
        self.segments = None        # ... per concatenated tower: [(first row, (c,h,w)), ...]

    # reference layout <-> device layout
    def to_device_layout(self, v):
        v = np.asarray(v, np.float32)
        if self.var.kind == 'convW':                   # (O,I,kh,kw) -> KC [(r,s,c)][o], flipped
            w = v[:, :, ::-1, ::-1].transpose(2, 3, 1, 0)
            return np.ascontiguousarray(w).reshape(-1)
        if self.var.kind == 'fcW' and self.chw is not None:
            c, h, w = self.chw
            return np.ascontiguousarray(v.reshape(c, h, w, -1).transpose(1, 2, 0, 3)).reshape(-1)
        if self.var.kind == 'fcW' and self.segments is not None:
            out = np.array(v, np.float32, copy=True)
            for r0, (c, h, w) in self.segments:
                n = c * h * w
                out[r0:r0 + n] = v[r0:r0 + n].reshape(c, h, w, -1).transpose(1, 2, 0, 3).reshape(n, -1)
            return out.reshape(-1)
        return np.ascontiguousarray(v).reshape(-1)

    def from_device_layout(self, flat):
        if self.var.kind == 'convW':
            o, i, kh, kw = self.shape
            w = flat.reshape(kh, kw, i, o).transpose(3, 2, 0, 1)[:, :, ::-1, ::-1]
            return np.ascontiguousarray(w)
        if self.var.kind == 'fcW' and self.chw is not None:
            c, h, w = self.chw
            return np.ascontiguousarray(flat.reshape(h, w, c, -1).transpose(2, 0, 1, 3)).reshape(self.shape)
        if self.var.kind == 'fcW' and self.segments is not None:
            v = flat.reshape(self.shape)
            out = v.copy()
            for r0, (c, h, w) in self.segments:
                n = c * h * w
                out[r0:r0 + n] = v[r0:r0 + n].reshape(h, w, c, -1).transpose(2, 0, 1, 3).reshape(n, -1)
            return out
        return flat.reshape(self.shape).copy()


class Engine(object):
    def __init__(self, net, precision=None, device=None, batch=None):
        """``batch``: samples per step on THIS device.  Defaults to the net's batch size; a data-parallel rank
        passes its shard of the global minibatch (trainer/nettrainer.py: global batch / world)."""
        torch = _torch()
        if not torch.cuda.is_available():
            raise DppError("dpp_b200.Engine needs a CUDA device (sm_100a); there is no CPU fallback")
        lib.load()
        lib.dpp_wgrad_workspace_init()     # library-owned scratch: allocated here, never inside a graph capture
        lib.dpp_fc_workspace_init()
        self.torch = torch
        self.dev = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        self.net = net
        self.output_sym = net.output
        self.B = int(net.cfgParams.batch_size) if batch is None else int(batch)
        if self.B < 1:
            raise DppError("Engine batch must be >= 1")
        if precision is None:
            precision = int(os.environ.get('DPP_PRECISION', '1'))
        self.precision = precision
        self.world = 1
        self.rank = 0
        self.allreduce_fn = None
        self.syncbn = False
        self._peer = None
        self._graphs = {}
        self._masks_injected = None
        # backward-weights kernels run on a second stream (forked/joined inside the step, also under graph
        # capture): their results are only needed by ADAM, so they fill the SMs the dgrad chain leaves idle
        self._wgrad_stream = torch.cuda.Stream(device=self.dev) if os.environ.get('DPP_WGRAD_STREAM', '1') != '0' else None
        # data-parallel runs: the FC tail's gradients (90 % of the arena, complete when the backward pass has
        # barely started) are all-reduced on this stream underneath the whole conv backward
        self._comm_stream = None
        self._lower()
        self._alloc_params()
        self._alloc_activations()
        self._train_ready = False
        self._alloc_packs()

    # ---------------------------------------------------------------------------------
    # lowering
    # ---------------------------------------------------------------------------------
    def _lower(self):
        from net.convlayer import ConvLayer
        from net.convpoollayer import ConvPoolLayer
        from net.batchnormlayer import BatchNormLayer
        from net.nonlinearitylayer import NonlinearityLayer
        from net.hiddenlayer import HiddenLayer
        from net.dropoutlayer import DropoutLayer

        out = self.output_sym
        order, seen = [], set()

        def visit(s):
            if id(s) in seen:
                return
            seen.add(id(s))
            for i in s.inputs:
                visit(i)
            order.append(s)
        visit(out)
        consumers = {}
        for s in order:
            for i in s.inputs:
                consumers.setdefault(id(i), []).append(s)
        pos = {id(s): k for k, s in enumerate(order)}

        def lay(s, cls):
            return s.op == 'layer' and isinstance(s.layer, cls)

        self.tensors = []
        self.ops = []
        self.bns = []                 # BatchNormLayer objects in use, in forward order
        val = {}                      # id(sym) -> Tensor
        virt = {}                     # id(relu sym) -> (bn layer, raw Tensor) not materialised
        inputs = [s for s in order if s.op == 'input']
        if isinstance(self.net.inputVar, (list, tuple)):
            # multi-input net (ScaleNet): inputs in the order of net.inputVar, dims from cfgParams.inputDim[k]
            inputs = [v for v in self.net.inputVar if id(v) in seen]
            dims = list(self.net.cfgParams.inputDim)
        else:
            if len(inputs) != 1:
                raise NotImplementedError("a single-input net with %d graph inputs" % len(inputs))
            dims = [self.net.cfgParams.inputDim]
        self.t_ins = []
        for k, (sym_in, dim) in enumerate(zip(inputs, dims)):
            tin = Tensor('input' if k == 0 else 'input%d' % k, dim, self.B)
            tin.is_input = True
            self.tensors.append(tin)
            self.t_ins.append(tin)
            val[id(sym_in)] = tin
        self.t_in = self.t_ins[0]

        def new_tensor(name, shape):
            t = Tensor(name, shape, self.B)
            self.tensors.append(t)
            return t

        # which conv fuses which add
        fused_add = {}      # id(conv sym) -> (add sym, residual operand sym)
        for s in order:
            if s.op == 'add':
                a, b = s.inputs
                cands = [x for x in (a, b) if lay(x, ConvLayer) and len(consumers[id(x)]) == 1]
                if not cands:
                    raise NotImplementedError("add without a ConvLayer operand")
                conv = max(cands, key=lambda x: pos[id(x)])
                other = b if conv is a else a
                fused_add[id(conv)] = (s, other)

        self.dropout_layers = []
        for s in order:
            if s.op == 'input':
                continue
            if s.op == 'add':
                continue                      # value assigned by the fusing conv
            if s.op == 'flatten':
                src = s.inputs[0]
                if id(src) in virt:           # BN->ReLU->flatten: materialise
                    bn, raw = virt[id(src)]
                    t = new_tensor('bnrelu%d' % bn.layerNum, raw.shape)
                    self.ops.append(dict(kind='bn_apply', bn=bn, src=raw, dst=t, relu=1))
                    val[id(src)] = t
                t4 = val[id(src)]
                t2 = Tensor(t4.name + '_flat', (t4.shape[0], int(np.prod(t4.shape[1:]))))
                t2.alias = t4
                t2.chw = t4.shape[1:]
                val[id(s)] = t2
                continue
            if s.op == 'concat':
                srcs = [val[id(i)] for i in s.inputs]
                if not all(isinstance(t, Tensor) and len(t.shape) == 2 for t in srcs):
                    raise NotImplementedError("concat of non-flattened tensors")
                cat = new_tensor('concat', (srcs[0].shape[0], int(sum(t.shape[1] for t in srcs))))
                cat.segments, off = [], 0
                for t in srcs:
                    if t.chw is not None:
                        cat.segments.append((off, tuple(t.chw)))
                    off += t.shape[1]
                self.ops.append(dict(kind='concat', srcs=srcs, dst=cat))
                val[id(s)] = cat
                continue
            if s.op == 'reshape':
                raise NotImplementedError("reshape hidden->conv is unused on the hot path")
            L = s.layer
            src = s.inputs[0]
            if isinstance(L, BatchNormLayer):
                raw = val[id(src)]
                if raw.bn is not None and raw.bn is not L:
                    raise NotImplementedError("two BN layers on one tensor")
                raw.bn = L
                self.bns.append(L)
                val[id(s)] = ('bn', L, raw)
                continue
            if isinstance(L, NonlinearityLayer):
                v = val[id(src)]
                if not (isinstance(v, tuple) and v[0] == 'bn'):
                    raise NotImplementedError("ReLU layer without a preceding BN")
                virt[id(s)] = (v[1], v[2])
                continue
            if isinstance(L, ConvLayer):
                p = L.cfgParams
                if id(src) in virt:
                    bn, raw = virt[id(src)]
                    xin, in_bn = raw, bn
                else:
                    xin, in_bn = val[id(src)], None
                t = new_tensor('conv%d' % L.layerNum, p.outputDim)
                op = dict(kind='conv', layer=L, src=xin, in_bn=in_bn, dst=t, residual=None)
                self.ops.append(op)
                val[id(s)] = t
                if id(s) in fused_add:
                    add_sym, other = fused_add[id(s)]
                    op['residual'] = val[id(other)]
                    val[id(add_sym)] = t
                continue
            if isinstance(L, ConvPoolLayer):
                p = L.cfgParams
                xin = val[id(src)]
                t = new_tensor('convpool%d' % L.layerNum, p.outputDim)
                self.ops.append(dict(kind='convpool', layer=L, src=xin, dst=t))
                val[id(s)] = t
                continue
            if isinstance(L, HiddenLayer):
                xin = val[id(src)]
                t = new_tensor('fc%d' % L.layerNum, (self.B, int(L.cfgParams.outputDim[1])))
                op = dict(kind='fc', layer=L, src=xin, dst=t, dropout=None)
                self.ops.append(op)
                val[id(s)] = t
                continue
            if isinstance(L, DropoutLayer):
                xin = val[id(src)]
                prod = [o for o in self.ops if o['kind'] == 'fc' and o['dst'] is xin]
                if not prod or len(consumers[id(src)]) != 1:
                    raise NotImplementedError("Dropout must directly follow a HiddenLayer")
                prod[0]['dropout'] = L
                self.dropout_layers.append(L)
                val[id(s)] = xin
                continue
            raise NotImplementedError("layer %r" % L)
        self.t_out = val[id(out)]
        if not isinstance(self.t_out, Tensor):
            raise NotImplementedError("network output must be a materialised tensor")
        # stats producers: every tensor with a BN must be produced by conv / convpool
        for op in self.ops:
            if op['kind'] in ('conv', 'convpool') and op['dst'].bn is not None:
                op['out_bn'] = op['dst'].bn
            else:
                op['out_bn'] = None
        for t in self.tensors:
            if t.bn is not None and not any(o.get('out_bn') is t.bn for o in self.ops):
                raise NotImplementedError("BN on a tensor not produced by a conv layer")
        # consumers of each BN's normalised output (for the backward accumulation order)
        self.bn_consumers = {}
        for op in self.ops:
            if op['kind'] == 'conv' and op['in_bn'] is not None:
                self.bn_consumers.setdefault(id(op['in_bn']), []).append(op)
            if op['kind'] == 'bn_apply':
                self.bn_consumers.setdefault(id(op['bn']), []).append(op)

    # ---------------------------------------------------------------------------------
    # parameters
    # ---------------------------------------------------------------------------------
    def _alloc_params(self):
        torch = self.torch
        self.slots = {}
        off = 0
        order = []
        for p in self.net.all_params:
            s = Slot(p, 'w', off)
            self.slots[id(p)] = s
            order.append(s)
            off += (s.size + 3) // 4 * 4          # 16-byte aligned slots
        self.n_w = off
        self.w_slots = order
        roff = 0
        self.r_slots = []
        for l in self.net.layers:
            for p in l.params_nontrained:
                s = Slot(p, 'r', roff)
                self.slots[id(p)] = s
                self.r_slots.append(s)
                roff += (s.size + 3) // 4 * 4
        self.n_r = max(roff, 4)
        # FC layers fed by a flattened NHWC tensor: permute rows
        for op in self.ops:
            if op['kind'] == 'fc' and getattr(op['src'], 'chw', None) is not None:
                self.slots[id(op['layer'].W)].chw = tuple(op['src'].chw)
            if op['kind'] == 'fc' and getattr(op['src'], 'segments', None):
                self.slots[id(op['layer'].W)].segments = list(op['src'].segments)
        self.W = torch.zeros(self.n_w, dtype=torch.float32, device=self.dev)
        self.R = torch.zeros(self.n_r, dtype=torch.float32, device=self.dev)
        self.G = self.M = self.V = None
        host = np.zeros(self.n_w, np.float32)
        for s in order:
            host[s.offset:s.offset + s.size] = s.to_device_layout(s.var._host)
        self.W.copy_(torch.from_numpy(host))
        hostr = np.zeros(self.n_r, np.float32)
        for s in self.r_slots:
            hostr[s.offset:s.offset + s.size] = s.to_device_layout(s.var._host)
        self.R.copy_(torch.from_numpy(hostr))
        for s in list(order) + self.r_slots:
            s.var._binding = (self, s)
        # BN statistic arena (fp64): per BN forward sums [2C] + backward sums [2C]
        soff = 0
        self.bn_stat_off = {}
        for bn in self.bns:
            c = int(bn.cfgParams.inputDim[1])
            self.bn_stat_off[id(bn)] = (soff, soff + 2 * c, c)
            soff += 4 * c
        self.n_stats = max(soff, 2)
        self.STATS = torch.zeros(self.n_stats, dtype=torch.float64, device=self.dev)
        # two 32-bit words per BN for the grid-wide barrier of the fused dgrad + BN-backward kernel (zeroed per step)
        self.GBAR = torch.zeros(2 * max(len(self.bns), 1), dtype=torch.int32, device=self.dev)
        self.bn_index = {id(bn): i for i, bn in enumerate(self.bns)}
        self.fuse_bn_bwd = os.environ.get('DPP_FUSE_BN_BWD', '1') != '0'
        # the backward-weights GEMMs of all ConvLayers in a few persistent launches at the end of the backward pass
        # (dpp_wgrad_group_*) instead of one launch per layer on a second stream
        self.group_wgrad = self.precision != 0 and os.environ.get('DPP_WGRAD_GROUP', '1') != '0'
        self._wgroup = None
        self._n_bn_bwd_launches = None

    def release(self):
        """Detach variables from the device arenas (values are pulled back to the host)."""
        for s in list(self.w_slots) + self.r_slots:
            if s.var._binding is not None and s.var._binding[0] is self:
                s.var._host = self.download_param(s)
                s.var._binding = None
        self._graphs = {}
        if getattr(self, '_wgroup', None) is not None:
            self.torch.cuda.synchronize()
            lib.dpp_wgrad_group_destroy(self._wgroup)
            self._wgroup = None
        if getattr(self, '_peer', None) is not None:
            self.torch.cuda.synchronize()
            for p in self._peer['opened']:
                lib.dpp_peer_close(p)
            lib.dpp_peer_free(self._peer['local'])
            self._peer = None

    def _arena(self, slot):
        return self.W if slot.arena == 'w' else self.R

    def download_param(self, slot):
        flat = self._arena(slot)[slot.offset:slot.offset + slot.size].cpu().numpy()
        return slot.from_device_layout(flat)

    def upload_param(self, slot, value):
        flat = self.torch.from_numpy(slot.to_device_layout(value))
        self._arena(slot)[slot.offset:slot.offset + slot.size].copy_(flat)
        if slot.var.kind == 'convW':
            self._packs_dirty = True

    # ---------------------------------------------------------------------------------
    # tcgen05 weight images (precision 1|2): re-packed after every optimiser step
    # ---------------------------------------------------------------------------------
    def _alloc_packs(self):
        torch = self.torch
        self._packs_dirty = False
        self.pack_items = None
        convs = [op for op in self.ops if op['kind'] == 'conv']
        if self.precision == 0 or not convs:
            return
        passes = 2 if self.precision == 1 else 1
        sizes, total = [], 0
        for op in convs:
            p = op['layer'].cfgParams
            cin, cout, k = int(p.inputDim[1]), int(p.nFilters), int(p.filterDim[0])
            f, g = C.c_int64(), C.c_int64()
            lib.dpp_conv_pack_size(cin, cout, k, self.precision, C.byref(f), C.byref(g))
            sizes.append((total, total + f.value, cin, cout, k))
            total += f.value + g.value
        self.WPACK = torch.zeros(total, dtype=torch.float32, device=self.dev)
        items = (PackItem * len(convs))()
        for i, (op, (o_f, o_g, cin, cout, k)) in enumerate(zip(convs, sizes)):
            op['wpack_fwd'] = self.WPACK.data_ptr() + 4 * o_f
            op['wpack_dgrad'] = self.WPACK.data_ptr() + 4 * o_g
            items[i].w = self.pview(op['layer'].W).data_ptr()
            items[i].img_fwd = op['wpack_fwd']
            items[i].img_dgrad = op['wpack_dgrad']
            items[i].Cin, items[i].Cout, items[i].k = cin, cout, k
            items[i].bn_fwd = min(cout, 128)
            items[i].bn_dgrad = min(cin, 128)
            items[i].passes = passes
        self.pack_items = torch.frombuffer(bytearray(bytes(items)), dtype=torch.uint8).to(self.dev)
        self.n_pack = len(convs)
        self._packs_dirty = True

    def _repack(self):
        if self.pack_items is not None:
            lib.dpp_conv_pack_all(_ptr(self.pack_items), self.n_pack, self._stream())
        self._packs_dirty = False

    def pview(self, var, arena=None):
        s = self.slots[id(var)]
        a = self._arena(s) if arena is None else arena
        return a[s.offset:s.offset + s.size]

    # ---------------------------------------------------------------------------------
    # activations
    # ---------------------------------------------------------------------------------
    def _nhwc_shape(self, t):
        s = t.shape
        return (s[0], s[2], s[3], s[1]) if len(s) == 4 else s

    def _alloc_activations(self):
        torch = self.torch
        for t in self.tensors:
            t.buf = torch.zeros(self._nhwc_shape(t), dtype=torch.float32, device=self.dev)
        for op in self.ops:
            if op['kind'] == 'convpool':
                op['argmax'] = torch.zeros(self._nhwc_shape(op['dst']), dtype=torch.uint8, device=self.dev)
        self.x_nchw_all = [torch.zeros(t.shape, dtype=torch.float32, device=self.dev) for t in self.t_ins]
        self.x_nchw = self.x_nchw_all[0]
        self.cost = torch.zeros(1, dtype=torch.float32, device=self.dev)

    def _alloc_training(self):
        if self._train_ready:
            return
        torch = self.torch
        self.G = torch.zeros_like(self.W)
        self.M = torch.zeros_like(self.W)
        self.V = torch.zeros_like(self.W)
        for t in self.tensors:
            if not t.is_input:
                t.grad = torch.zeros_like(t.buf)
        self.dz = {}
        for bn in self.bns:
            raw = [t for t in self.tensors if t.bn is bn][0]
            self.dz[id(bn)] = torch.zeros_like(raw.buf)
        for op in self.ops:
            if op['kind'] == 'fc':
                op['scratch'] = torch.zeros_like(op['dst'].buf)
                if op['dropout'] is not None:
                    op['mask'] = torch.ones_like(op['dst'].buf)
                    g = torch.Generator(device=self.dev)
                    g.manual_seed(op['dropout'].mask_seed)
                    op['mask_gen'] = g
        self.y_in = torch.zeros((self.B, int(np.prod(self.t_out.shape[1:]))), dtype=torch.float32, device=self.dev)
        # hyper: lr, t, -, grad_scale
        self.hyper = torch.tensor([0.0, 1.0, 0.0, 1.0], dtype=torch.float32, device=self.dev)
        self.hyper_host = torch.zeros(4, dtype=torch.float32).pin_memory()
        if getattr(self.net, 'params_filter', None):
            raise NotImplementedError("net.params_filter (frozen layers): the ADAM kernel updates the whole arena")
        alphas = set(float(bn.cfgParams.alpha) for bn in self.bns)
        if len(alphas) > 1:
            raise NotImplementedError("BatchNorm layers with different EMA factors: %s" % sorted(alphas))
        self._build_ema_items()
        self._train_ready = True

    def _build_ema_items(self):
        """table of the running-statistics update (one launch for all BNs); the sample count behind the fp64 sums
        is the global one under SyncBN"""
        torch = self.torch
        items = (BnEmaItem * max(len(self.bns), 1))()
        for i, bn in enumerate(self.bns):
            f0, _, c = self.bn_stat_off[id(bn)]
            raw = [t for t in self.tensors if t.bn is bn][0]
            items[i].sums = self.STATS.data_ptr() + 8 * f0
            items[i].mean = self.pview(bn.mean).data_ptr()
            items[i].inv_std = self.pview(bn.inv_std).data_ptr()
            items[i].count = float(raw.pixels) * (self.world if self.syncbn else 1)
            items[i].C = c
            items[i].eps = float(bn.cfgParams.epsilon)
        raw_bytes = bytes(items)
        self.ema_items = torch.frombuffer(bytearray(raw_bytes), dtype=torch.uint8).to(self.dev)

    # ---------------------------------------------------------------------------------
    # kernels
    # ---------------------------------------------------------------------------------
    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def _bnref(self, bn, raw, train, relu=1):
        r = BnRef()
        f0, _, c = self.bn_stat_off[id(bn)]
        r.sums = (self.STATS.data_ptr() + 8 * f0) if train else None
        r.mean = self.pview(bn.mean).data_ptr()
        r.inv_std = self.pview(bn.inv_std).data_ptr()
        r.gamma = self.pview(bn.gamma).data_ptr()
        r.beta = self.pview(bn.beta).data_ptr()
        r.count = float(raw.pixels) * (self.world if self.syncbn else 1)
        r.eps = float(bn.cfgParams.epsilon)
        r.relu = relu
        return r

    def _conv_desc(self, op):
        L = op['layer']
        p = L.cfgParams
        n, ci, h, w = p.inputDim
        d = ConvDesc()
        d.N, d.H, d.W, d.Cin = self.B, int(h), int(w), int(ci)
        d.Cout = int(p.nFilters)
        d.k = int(p.filterDim[0])
        d.stride = int(p.stride[0])
        if p.border_mode != 'half':
            raise NotImplementedError("ConvLayer border_mode %s" % p.border_mode)
        d.pad = d.k // 2
        d.Ho, d.Wo = int(p.outputDim[2]), int(p.outputDim[3])
        d.precision = self.precision if 'wpack_fwd' in op else 0
        d.wpack_fwd = op.get('wpack_fwd')
        d.wpack_dgrad = op.get('wpack_dgrad')
        return d

    def _stats_ptr(self, bn, which=0):
        f0, b0, _ = self.bn_stat_off[id(bn)]
        return C.c_void_p(self.STATS.data_ptr() + 8 * (f0 if which == 0 else b0))

    def _sync_stats(self, bn, which):
        """SyncBN: sum this BN's fp64 statistics ({sum, sumsq} forward / {sum dz, sum dz*xhat} backward) over the
        data-parallel ranks, in stream order between the kernel that produced them and the one that reads them, so
        that G ranks x B/G samples normalise exactly like one device with B samples (net/batchnormlayer.py:154-159
        takes the statistics over the whole minibatch)."""
        f0, b0, c = self.bn_stat_off[id(bn)]
        lo = f0 if which == 0 else b0
        if self._peer is not None:
            # one-shot exchange over peer memory (NVLink P2P): one small kernel in stream order instead of a collective
            pr = self._peer
            lib.dpp_stats_exchange(C.c_void_p(self.STATS.data_ptr() + 8 * lo), 2 * c, _ptr(pr['ptrs']),
                                   pr['regions'][(id(bn), which)], self.rank, self.world, _ptr(pr['seq']), _ptr(pr['err']),
                                   self._stream())
        else:
            self.stats_allreduce_fn(self.STATS[lo:lo + 2 * c])

    def _setup_peer_exchange(self, dist):
        """exchange buffers of the SyncBN statistics, mapped into every rank over CUDA IPC (same node, NVLink)"""
        torch = self.torch
        regions, off = {}, 0
        for bn in self.bns:
            c = self.bn_stat_off[id(bn)][2]
            for which in (0, 1):
                regions[(id(bn), which)] = off
                off += ((self.world * (2 * c + 1) * 8 + 15) // 16) * 16
        local, handle = C.c_void_p(), C.create_string_buffer(64)
        lib.dpp_peer_alloc(max(off, 16), C.byref(local), handle)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(handle.raw))
        ptrs, opened = [], []
        for r in range(self.world):
            if r == self.rank:
                ptrs.append(local.value)
            else:
                p = C.c_void_p()
                lib.dpp_peer_open(handles[r], C.byref(p))
                ptrs.append(p.value)
                opened.append(p)
        self._peer = dict(regions=regions, local=local, opened=opened,
                          ptrs=torch.tensor(ptrs, dtype=torch.int64, device=self.dev),
                          seq=torch.zeros(1, dtype=torch.int64, device=self.dev),
                          err=torch.zeros(1, dtype=torch.int32, device=self.dev))
        dist.barrier()          # every rank has mapped every buffer before the first exchange is issued

    def _run_forward(self, train):
        st = self._stream()
        for op in self.ops:
            k = op['kind']
            if k == 'convpool':
                L = op['layer']
                p = L.cfgParams
                n, ci, h, w = p.inputDim
                pad = p.filterDim[0] // 2 if p.border_mode == 'half' else 0
                if p.stride != (1, 1) and tuple(p.stride) != (1, 1):
                    raise NotImplementedError("strided ConvPoolLayer")
                relu = 1 if p.activation_str == 'ReLU' else 0
                stats = self._stats_ptr(op['out_bn']) if (op['out_bn'] is not None and train) else None
                lib.dpp_convpool_fwd(_ptr(op['src'].buf), _ptr(self.pview(L.W)), _ptr(self.pview(L.b)),
                                     _ptr(op['dst'].buf), _ptr(op['argmax']), stats, self.B, int(h), int(w), int(ci),
                                     int(p.nFilters), int(p.filterDim[0]), int(pad), int(p.poolsize[0]), relu, st)
            elif k == 'conv':
                L = op['layer']
                d = self._conv_desc(op)
                bnref = self._bnref(op['in_bn'], op['src'], train) if op['in_bn'] is not None else None
                stats = self._stats_ptr(op['out_bn']) if (op['out_bn'] is not None and train) else None
                lib.dpp_conv2d_fwd(C.byref(d), _ptr(op['src'].buf), C.byref(bnref) if bnref else None,
                                   _ptr(self.pview(L.W)), _ptr(self.pview(L.b)),
                                   _ptr(op['residual'].buf) if op['residual'] is not None else None,
                                   _ptr(op['dst'].buf), stats, st)
            elif k == 'concat':
                dst = op['dst'].buf
                off = 0
                for t in op['srcs']:
                    buf = t.alias.buf if hasattr(t, 'alias') else t.buf
                    n = int(t.shape[1])
                    lib.dpp_copy2d(C.c_void_p(dst.data_ptr() + 4 * off), 4 * int(dst.shape[1]), _ptr(buf), 4 * n, 4 * n,
                                   self.B, st)
                    off += n
            elif k == 'bn_apply':
                bnref = self._bnref(op['bn'], op['src'], train, relu=op['relu'])
                c = op['src'].shape[1]
                lib.dpp_bn_apply(_ptr(op['src'].buf), C.byref(bnref), _ptr(op['dst'].buf), op['src'].pixels, int(c), st)
            elif k == 'fc':
                L = op['layer']
                src = op['src']
                xbuf = src.alias.buf if hasattr(src, 'alias') else src.buf
                n_in, n_out = int(L.cfgParams.inputDim[1]), int(L.cfgParams.outputDim[1])
                relu = 1 if L.cfgParams.activation_str == 'ReLU' else 0
                mask, scale = None, 1.0
                if op['dropout'] is not None:
                    if train:
                        mask = op['mask']
                    else:
                        scale = float(op['dropout'].prob_keep)
                lib.dpp_fc_fwd(_ptr(xbuf), _ptr(self.pview(L.W)), _ptr(self.pview(L.b)), _ptr(op['dst'].buf),
                               self.B, n_in, n_out, relu, _ptr(mask), scale, self.precision, st)
            else:
                raise NotImplementedError(k)
            if train and self.syncbn and op.get('out_bn') is not None:
                self._sync_stats(op['out_bn'], 0)       # the consumer's prologue reads the global sums

    def _grad_slots_of(self, op, bn_done):
        """arena slots whose gradients are issued by the backward kernels of ``op`` (+ the BNs it completes)"""
        out = []
        if op['kind'] in ('convpool', 'fc') or (op['kind'] == 'conv' and not self.group_wgrad):
            L = op['layer']
            out += [self.slots[id(L.W)], self.slots[id(L.b)]]     # (grouped backward-weights: complete only at the very end)
        for bn in bn_done:
            out += [self.slots[id(bn.gamma)], self.slots[id(bn.beta)]]
        return out

    def _plan_buckets(self, min_elems):
        """Data-parallel exchange plan (dp.plan_buckets): which slices of the gradient arena can be summed over the
        ranks after which op of the reverse walk.  Returns ({index in reversed(self.ops): (lo, hi)}, trailing)."""
        from . import dp
        pending = {id(bn): len(self.bn_consumers.get(id(bn), [])) for bn in self.bns}
        events, rev = [], list(reversed(self.ops))
        for i, op in enumerate(rev):
            finished = []
            bn = op.get('in_bn') if op['kind'] == 'conv' else (op.get('bn') if op['kind'] == 'bn_apply' else None)
            if bn is not None:
                pending[id(bn)] -= 1
                if pending[id(bn)] == 0:
                    finished.append(bn)
            force = op['kind'] == 'fc' and (i + 1 == len(rev) or rev[i + 1]['kind'] != 'fc')    # end of the FC tail
            events.append(([sl.offset for sl in self._grad_slots_of(op, finished)], force))
        return dp.plan_buckets([sl.offset for sl in self.w_slots], self.n_w, events, min_elems)

    def _run_backward(self):
        """Reverse walk.  t.grad of the output tensor must hold dCost/dOut on entry."""
        st = self._stream()
        G = self.G
        torch = self.torch
        main = torch.cuda.current_stream()
        side = self._wgrad_stream
        side_ptr = C.c_void_p(side.cuda_stream) if side is not None else None
        forked = False
        pending = {}          # id(bn) -> number of consumers still to contribute
        for bn in self.bns:
            pending[id(bn)] = len(self.bn_consumers.get(id(bn), []))
        # a tensor used as residual operand receives the fusing conv's output gradient as `skip`
        skip_of = {}
        for op in self.ops:
            if op['kind'] == 'conv' and op['residual'] is not None:
                skip_of[id(op['residual'])] = op['dst']
        # parameter gradients of a BN are the global sums under SyncBN: every rank adds 1/world of them, so that the
        # SUM all-reduce of the arena (and ADAM's 1/world) yields them exactly once
        pscale = (1.0 / self.world) if self.syncbn else 1.0

        def finish_bn(bn, raw):
            """all consumers contributed dz (+stats): apply the BN backward, producing raw.grad"""
            if self.syncbn:
                self._sync_stats(bn, 1)
            bnref = self._bnref(bn, raw, True)
            c = raw.shape[1]
            sk = skip_of.get(id(raw))
            lib.dpp_bn_bwd_apply(_ptr(self.dz[id(bn)]), _ptr(raw.buf), C.byref(bnref), self._stats_ptr(bn, 1),
                                 _ptr(sk.grad) if sk is not None else None, _ptr(raw.grad),
                                 _ptr(self.pview(bn.gamma, G)), _ptr(self.pview(bn.beta, G)), None,
                                 raw.pixels, int(c), pscale, st)

        # data-parallel exchange: buckets of the gradient arena, all-reduced on the comm stream as soon as the
        # backward kernels that fill them have been issued (on the main stream and on the backward-weights stream)
        cuts, trailing = {}, None
        exchanging = self.allreduce_fn is not None and self.world > 1
        if exchanging:
            if os.environ.get('DPP_EARLY_ALLREDUCE', '1') != '0':
                cuts, trailing = self._bucket_plan
            else:
                trailing = (0, self.n_w)
            if self._comm_stream is None:
                self._comm_stream = torch.cuda.Stream(device=self.dev)
        self._exchanged = []
        n_fused = [0]
        wlayers = []

        def exchange(lo, hi):
            comm = self._comm_stream
            comm.wait_stream(main)
            if side is not None and forked:
                comm.wait_stream(side)
            with torch.cuda.stream(comm):
                self.allreduce_fn(G[lo:hi])
            self._exchanged.append((lo, hi))

        for i, op in enumerate(reversed(self.ops)):
            k = op['kind']
            if k == 'fc':
                L = op['layer']
                src = op['src']
                base = src.alias if hasattr(src, 'alias') else src
                n_in, n_out = int(L.cfgParams.inputDim[1]), int(L.cfgParams.outputDim[1])
                relu = 1 if L.cfgParams.activation_str == 'ReLU' else 0
                mask = op['mask'] if op['dropout'] is not None else None
                dx = None if base.is_input else base.grad
                # dW is ASSIGNED (flag 1 = DPP_FC_DW_ASSIGN): _step_body's zero fill skips the HiddenLayer weight slots
                lib.dpp_fc_bwd_ex(_ptr(base.buf), _ptr(self.pview(L.W)), _ptr(op['dst'].buf), _ptr(op['dst'].grad),
                                  _ptr(self.pview(L.W, G)), _ptr(self.pview(L.b, G)), _ptr(dx), _ptr(op['scratch']),
                                  self.B, n_in, n_out, relu, _ptr(mask), 1.0, self.precision, 1, st)
            elif k == 'concat':
                raise NotImplementedError("backward through a tower concatenation: ScaleNet training is out of scope "
                                          "(DESIGN.md section 8)")
            elif k == 'bn_apply':
                bn, raw = op['bn'], op['src']
                bnref = self._bnref(bn, raw, True, relu=op['relu'])
                c = raw.shape[1]
                lib.dpp_bn_relu_bwd_reduce(_ptr(op['dst'].grad), _ptr(raw.buf), C.byref(bnref), _ptr(self.dz[id(bn)]),
                                           self._stats_ptr(bn, 1), raw.pixels, int(c), st)
                pending[id(bn)] -= 1
                if pending[id(bn)] == 0:
                    finish_bn(bn, raw)
            elif k == 'conv':
                L = op['layer']
                d = self._conv_desc(op)
                dst = op['dst']
                # a conv output that is only the residual operand of a later conv (projection
                # block: conv3 feeding the shortcut conv's epilogue) shares that conv's gradient
                dy = skip_of[id(dst)].grad if (id(dst) in skip_of and dst.bn is None) else dst.grad
                bn, raw = op['in_bn'], op['src']
                bnref = self._bnref(bn, raw, True) if bn is not None else None
                if self.group_wgrad:
                    if self._wgroup is None:
                        wl = WgradLayer()
                        wl.d = d
                        wl.x = raw.buf.data_ptr()
                        if bnref is not None:
                            wl.in_bn, wl.has_in_bn = bnref, 1
                        wl.dy = dy.data_ptr()
                        wl.dw = self.pview(L.W, G).data_ptr()
                        wl.db = self.pview(L.b, G).data_ptr()
                        wlayers.append(wl)
                else:
                    if side is not None:
                        side.wait_stream(main)          # dy (and everything before it) is complete
                        forked = True
                    lib.dpp_conv2d_wgrad(C.byref(d), _ptr(raw.buf), C.byref(bnref) if bnref else None, _ptr(dy),
                                         _ptr(self.pview(L.W, G)), _ptr(self.pview(L.b, G)), side_ptr if side is not None else st)
                if bn is not None:
                    total = len(self.bn_consumers[id(bn)])
                    first = (pending[id(bn)] == total)
                    last = (pending[id(bn)] == 1)
                    dzbuf = self.dz[id(bn)]
                    if first and d.stride != 1:
                        lib.dpp_fill_f32(_ptr(dzbuf), 0.0, dzbuf.numel(), st)
                    fused = False
                    if last and self.fuse_bn_bwd and not self.syncbn and d.stride == 1 and not op.get('no_tail'):
                        # the LAST dgrad into a normalised tensor also applies that BatchNorm's backward, behind a
                        # grid-wide barrier: one launch instead of two, dz re-read from L2
                        sk = skip_of.get(id(raw))
                        gb = C.c_void_p(self.GBAR.data_ptr() + 8 * self.bn_index[id(bn)])
                        rc = lib.raw('dpp_conv2d_dgrad_bn_bwd')(
                            C.byref(d), _ptr(dy), _ptr(self.pview(L.W)), _ptr(dzbuf), 0 if first else 1, C.byref(bnref),
                            _ptr(raw.buf), self._stats_ptr(bn, 1), _ptr(sk.grad) if sk is not None else None, _ptr(raw.grad),
                            _ptr(self.pview(bn.gamma, G)), _ptr(self.pview(bn.beta, G)), pscale, gb, st)
                        if rc == 0:
                            fused = True
                            n_fused[0] += 1
                        elif rc == DPP_ENOTSUP:
                            op['no_tail'] = True
                        else:
                            raise DppError("dpp_conv2d_dgrad_bn_bwd failed (%d): %s" % (rc, lib.raw('dpp_last_error')().decode()))
                    if not fused:
                        lib.dpp_conv2d_dgrad(C.byref(d), _ptr(dy), _ptr(self.pview(L.W)), _ptr(dzbuf), 0 if first else 1,
                                             C.byref(bnref) if last else None, _ptr(raw.buf) if last else None,
                                             self._stats_ptr(bn, 1) if last else None, st)
                    pending[id(bn)] -= 1
                    if last and not fused:
                        finish_bn(bn, raw)
                elif not raw.is_input:
                    lib.dpp_conv2d_dgrad(C.byref(d), _ptr(dy), _ptr(self.pview(L.W)), _ptr(raw.grad), 0, None, None,
                                         None, st)
            elif k == 'convpool':
                L = op['layer']
                p = L.cfgParams
                n, ci, h, w = p.inputDim
                pad = p.filterDim[0] // 2 if p.border_mode == 'half' else 0
                relu = 1 if p.activation_str == 'ReLU' else 0
                dx = None if op['src'].is_input else op['src'].grad
                lib.dpp_convpool_bwd(_ptr(op['src'].buf), _ptr(self.pview(L.W)), _ptr(op['dst'].buf),
                                     _ptr(op['argmax']), _ptr(op['dst'].grad), _ptr(self.pview(L.W, G)),
                                     _ptr(self.pview(L.b, G)), _ptr(dx), self.B, int(h), int(w), int(ci),
                                     int(p.nFilters), int(p.filterDim[0]), int(pad), int(p.poolsize[0]), relu, st)
            else:
                raise NotImplementedError(k)
            if i in cuts:
                exchange(*cuts[i])
        if self.group_wgrad:
            if self._wgroup is None and wlayers:
                if torch.cuda.is_current_stream_capturing():
                    raise DppError("the grouped backward-weights tables must be built outside graph capture "
                                   "(Engine.train_step runs one eager step first)")
                arr = (WgradLayer * len(wlayers))(*wlayers)
                h = C.c_void_p()
                rc = lib.raw('dpp_wgrad_group_create')(arr, len(wlayers), C.byref(h))
                if rc == DPP_ENOTSUP:
                    raise DppError("grouped backward-weights: a ConvLayer lies outside the tcgen05 path; set DPP_WGRAD_GROUP=0")
                if rc != 0:
                    raise DppError("dpp_wgrad_group_create failed (%d): %s" % (rc, lib.raw('dpp_last_error')().decode()))
                self._wgroup = h
            if self._wgroup is not None:
                lib.dpp_wgrad_group_run(self._wgroup, st)
        self._adam_tail = None
        if exchanging and trailing is not None:
            split = (cuts and trailing[0] == 0 and 0 < trailing[1] < self.n_w and trailing[1] % 4 == 0
                     and os.environ.get('DPP_SPLIT_ADAM', '1') != '0')
            if split:
                # everything above the trailing bucket has been handed to the comm stream already: the optimiser can
                # take [hi, n) while the last, small bucket (conv / BN gradients) is still being exchanged
                early_done = torch.cuda.Event()
                early_done.record(self._comm_stream)
            exchange(*trailing)
            if split:
                main.wait_event(early_done)
                self._adam_tail = trailing
        if forked:
            main.wait_stream(side)                  # join: the gradient arena is complete
        if exchanging and self._adam_tail is None:
            main.wait_stream(self._comm_stream)     # ... and summed over the ranks
        self._n_bn_bwd_launches = len(self.bns) - n_fused[0]

    def check_barriers(self):
        """raises if a grid-wide barrier of the fused dgrad + BN-backward kernels, or a wait of the peer-memory
        statistics exchange, gave up (synchronises)"""
        marks = self.GBAR[1::2].cpu().numpy()
        if (marks != 0).any():
            raise DppError("grid barrier timed out in %d fused BatchNorm-backward launch(es)" % int((marks != 0).sum()))
        if self._peer is not None and int(self._peer['err'].cpu()[0]) != 0:
            raise DppError("peer-memory statistics exchange timed out waiting for another rank")

    def launches_wgrad(self):
        """backward-weights launches per step: one per ConvLayer, or the grouped launches"""
        convs = len([o for o in self.ops if o['kind'] == 'conv'])
        if self.group_wgrad and self._wgroup is not None:
            return int(lib.dpp_wgrad_group_launches(self._wgroup))
        return convs

    def launches_bn_bwd(self):
        """separate BatchNorm-backward launches of the last step (the others ran fused behind their dgrad)"""
        return len(self.bns) if self._n_bn_bwd_launches is None else self._n_bn_bwd_launches

    # ---------------------------------------------------------------------------------
    # public API
    # ---------------------------------------------------------------------------------
    def forward_device(self, deterministic=True):
        """Run the forward pass on the NHWC input already in ``self.t_in.buf``; returns the device
        output tensor (B, n_out)."""
        st = self._stream()
        if self._packs_dirty:
            self._repack()
        if not deterministic:
            if self.dropout_layers:
                self._alloc_training()          # the masks live with the training buffers
                self._draw_masks()
            lib.dpp_fill_f64(_ptr(self.STATS), 0.0, self.STATS.numel(), st)
        self._run_forward(train=not deterministic)
        return self.t_out.buf

    def set_input_nchw(self, x_host_or_dev, which=0):
        torch = self.torch
        if isinstance(x_host_or_dev, np.ndarray):
            x_host_or_dev = torch.from_numpy(np.ascontiguousarray(x_host_or_dev, np.float32))
        stage, tin = self.x_nchw_all[which], self.t_ins[which]
        stage.copy_(x_host_or_dev.reshape(stage.shape), non_blocking=True)
        n, c, h, w = tin.shape
        lib.dpp_nchw_to_nhwc(_ptr(stage), _ptr(tin.buf), n, c, h, w, self._stream())

    def forward_host(self, batch_list, deterministic=True):
        """computeOutput's inner step: numpy NCHW batch(es) in, numpy output out."""
        if len(batch_list) != len(self.t_ins):
            raise DppError("network takes %d input(s), got %d" % (len(self.t_ins), len(batch_list)))
        for k, b in enumerate(batch_list):
            self.set_input_nchw(b, which=k)
        out = self.forward_device(deterministic=deterministic)
        return out.cpu().numpy()

    def set_dropout_masks(self, masks):
        """Inject explicit dropout masks (list of numpy (B, n) arrays, forward order) - parity
        tests only; production masks are drawn on the device each step."""
        self._alloc_training()
        self._masks_injected = True
        fcs = [o for o in self.ops if o['kind'] == 'fc' and o['dropout'] is not None]
        for o, m in zip(fcs, masks):
            o['mask'].copy_(self.torch.from_numpy(np.ascontiguousarray(m, np.float32)))

    def _draw_masks(self):
        if self._masks_injected:
            return
        for o in self.ops:
            if o['kind'] == 'fc' and o['dropout'] is not None:
                keep = float(o['dropout'].prob_keep)
                o['mask'].bernoulli_(keep, generator=o['mask_gen'])

    def _g_zero_ranges(self):
        """Ranges of the gradient arena that the backward kernels ACCUMULATE into.  The HiddenLayer weight gradients
        are assigned (dpp_fc_bwd_ex, DPP_FC_DW_ASSIGN) - for the ResNet that is 90 % of the arena (67 MB) that need
        not be cleared every step."""
        if getattr(self, '_g_zero', None) is None:
            skip = sorted((self.slots[id(o['layer'].W)].offset, self.slots[id(o['layer'].W)].size)
                          for o in self.ops if o['kind'] == 'fc')
            out, lo = [], 0
            for off, size in skip:
                if off > lo:
                    out.append((lo, off))
                lo = max(lo, off + size)
            if lo < self.n_w:
                out.append((lo, self.n_w))
            self._g_zero = out
        return self._g_zero

    def _step_body(self):
        """zero -> forward(train) -> cost -> backward -> (allreduce) -> ADAM -> EMA"""
        st = self._stream()
        lib.dpp_fill_f64(_ptr(self.STATS), 0.0, self.STATS.numel(), st)
        for lo, hi in self._g_zero_ranges():
            lib.dpp_fill_f32(_ptr(self.G[lo:hi]), 0.0, hi - lo, st)
        lib.dpp_fill_f32(_ptr(self.GBAR), 0.0, self.GBAR.numel(), st)      # all-zero words
        self._run_forward(train=True)
        d = int(self.y_in.shape[1])
        lib.dpp_loss_sqerr(_ptr(self.t_out.buf), _ptr(self.y_in), _ptr(self.t_out.grad), _ptr(self.cost), self.B, d, st)
        self._run_backward()
        tail = getattr(self, '_adam_tail', None)
        if tail is not None:
            # data parallel: ADAM over the part of the arena whose exchange is complete (the HiddenLayers: 90 % of the
            # parameters) hides the exchange of the trailing bucket; then the rest
            lo, hi = tail
            lib.dpp_adam_step(_ptr(self.W[hi:]), _ptr(self.G[hi:]), _ptr(self.M[hi:]), _ptr(self.V[hi:]), _ptr(self.hyper),
                              self.n_w - hi, st)
            self.torch.cuda.current_stream().wait_stream(self._comm_stream)
            lib.dpp_adam_step(_ptr(self.W[lo:hi]), _ptr(self.G[lo:hi]), _ptr(self.M[lo:hi]), _ptr(self.V[lo:hi]),
                              _ptr(self.hyper), hi - lo, st)
        else:
            lib.dpp_adam_step(_ptr(self.W), _ptr(self.G), _ptr(self.M), _ptr(self.V), _ptr(self.hyper), self.n_w, st)
        lib.dpp_adam_tick(_ptr(self.hyper), st)
        if self.pack_items is not None:
            lib.dpp_conv_pack_all(_ptr(self.pack_items), self.n_pack, st)
        if self.bns:
            lib.dpp_bn_ema_update(_ptr(self.ema_items), len(self.bns), float(self.bns[0].cfgParams.alpha), st)

    def set_lr(self, lr):
        self.hyper[0:1].fill_(float(lr))

    def set_world(self, world, allreduce_fn, rank=0, syncbn=False, stats_allreduce_fn=None, bucket_elems=None, dist=None):
        """Data-parallel mode (the reference is single-device).  ``allreduce_fn(t)`` sums a float32 slice
        of the gradient arena over the ranks in place, on the CURRENT stream (torch.distributed.all_reduce).
        ``syncbn``: also sum every BatchNorm's fp64 statistics (forward and backward) with ``stats_allreduce_fn``
        (default: the same function), which makes ``world`` ranks x B/world samples compute the single-device
        function of the global minibatch.  ``syncbn='p2p'`` (needs ``dist`` = torch.distributed, all ranks on one
        node) exchanges the statistics with dpp_stats_exchange over peer memory instead of a collective per BN."""
        self.world = int(world)
        self.rank = int(rank)
        self.allreduce_fn = allreduce_fn
        self.syncbn = bool(syncbn) and self.world > 1
        self.stats_allreduce_fn = stats_allreduce_fn or allreduce_fn
        if self.syncbn and syncbn == 'p2p' and self._peer is None:
            if dist is None:
                raise DppError("syncbn='p2p' needs dist=torch.distributed to exchange the IPC handles")
            self._setup_peer_exchange(dist)
        self._graphs = {}
        self._alloc_training()
        self._build_ema_items()                      # the counts behind the statistics depend on world / syncbn
        for op in self.ops:                          # every rank draws its own dropout masks
            if op['kind'] == 'fc' and op['dropout'] is not None:
                op['mask_gen'].manual_seed(op['dropout'].mask_seed + self.rank)
        self.hyper[3:4].fill_(1.0 / self.world)
        if bucket_elems is None:
            bucket_elems = int(os.environ.get('DPP_BUCKET_ELEMS', str(256 * 1024)))
        self._bucket_plan = self._plan_buckets(bucket_elems)

    def train_step(self, lr=None, use_graph=True):
        """One ``train_model`` call (poseregnettrainer.py:146-160) on the batch in
        ``t_in.buf`` (NHWC) / ``y_in``.  Returns the device scalar holding the minibatch cost."""
        self._alloc_training()
        torch = self.torch
        if lr is not None:
            self.set_lr(lr)
        if self._packs_dirty:
            self._repack()
        self._draw_masks()
        if not use_graph:
            self._step_body()
            return self.cost
        g = self._graphs.get('train')
        if g is None:
            # eager warm-up (sets kernel attributes, allocates nothing new), then capture
            w0, r0, m0, v0, h0 = self.W.clone(), self.R.clone(), self.M.clone(), self.V.clone(), self.hyper.clone()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._step_body()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            self.W.copy_(w0); self.R.copy_(r0); self.M.copy_(m0); self.V.copy_(v0); self.hyper.copy_(h0)
            self._repack()          # the warm-up step re-packed the post-ADAM weights: restore the images too
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            # thread_local: NCCL's watchdog thread may touch the CUDA API while we capture (world > 1)
            with torch.cuda.graph(g, capture_error_mode='thread_local' if self.world > 1 else 'global'):
                self._step_body()
            # the capture itself does not execute; state is untouched
            self._graphs['train'] = g
        g.replay()
        return self.cost

    def gradients(self):
        """dict var-id -> numpy gradient in the reference layout (parity tests)."""
        out = {}
        for s in self.w_slots:
            flat = self.G[s.offset:s.offset + s.size].cpu().numpy()
            out[id(s.var)] = s.from_device_layout(flat)
        return out

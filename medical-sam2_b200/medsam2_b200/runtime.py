"""Run-time settings shared by the host-side modules: the compute dtype of the GEMM/attention
operands (bf16 by default, fp32 = exact mode) and a cache of kernel-ready parameter copies
(compute-dtype weights, fp32 biases, re-laid-out conv weights) keyed on the parameter version so
that `load_state_dict` / in-place updates invalidate them."""
import contextlib
import os
import weakref

import torch

_COMPUTE = {"bf16": torch.bfloat16, "fp32": torch.float32}[os.environ.get("MS2_COMPUTE", "bf16")]


def compute_dtype():
    return _COMPUTE


def set_compute_dtype(dtype):
    global _COMPUTE
    assert dtype in (torch.float32, torch.bfloat16)
    _COMPUTE = dtype


@contextlib.contextmanager
def compute(dtype):
    global _COMPUTE
    old = _COMPUTE
    set_compute_dtype(dtype)
    try:
        yield
    finally:
        _COMPUTE = old


# Parameter generation: bumped whenever kernel-ready parameter copies may have been re-derived (a stale cache entry was
# rebuilt, `load_state_dict`, a train()/eval() switch, `.to()`/`_apply`).  Captured CUDA graphs bake in the device
# pointers of those copies, so GraphRunner drops every graph of an older generation instead of replaying stale or
# freed weights.
_GEN = [0]
_CAPTURE_REFS = None           # while a GraphRunner captures: the derived tensors the capture touched (kept alive with it)
_DERIVED_IN_CAPTURE = [False]  # a parameter copy had to be derived while capturing: its kernels were only RECORDED, so the
                               # value is never stored in the cache and the capture is thrown away (GraphRunner.run)


def bump_generation():
    _GEN[0] += 1


def generation():
    return _GEN[0]


class ParamCache:
    """derived tensors of parameters: get(param, key, fn) -> fn(param) cached per (version, ptr)."""

    def __init__(self):
        self._d = {}

    def get(self, params, key, fn):
        if isinstance(params, torch.Tensor):
            # fast path (one parameter): id -> (version, data_ptr, weakref, value); ~1 us on the hot path
            k = (id(params), key)
            hit = self._d.get(k)
            if hit is not None and hit[3] is not None:
                if hit[0][0] == params._version and hit[0][1] == params.data_ptr() and hit[2][0]() is params:
                    if _CAPTURE_REFS is not None:
                        _CAPTURE_REFS.append(hit[1])
                    return hit[1]
            if hit is not None:
                bump_generation()                     # a derived copy is being replaced: captured graphs point at the old one
            with torch.no_grad():
                val = fn(params)
            if _CAPTURE_REFS is not None:
                _DERIVED_IN_CAPTURE[0] = True         # recorded, not executed: never cache it
                _CAPTURE_REFS.append(val)
                return val
            if len(self._d) > 4096:
                self._d = {kk: v for kk, v in self._d.items() if all(r() is not None for r in v[2])}
            self._d[k] = ((params._version, params.data_ptr(), params.dtype, str(params.device)), val,
                          (weakref.ref(params),), True)
            return val
        sig = tuple((p.data_ptr(), p._version, p.dtype, str(p.device)) for p in params)
        k = (tuple(id(p) for p in params), key)
        hit = self._d.get(k)
        # id() and data_ptr() are both recycled once a model is freed: an entry is only valid while the very
        # tensor objects it was derived from are still alive (weak references), not merely "same address".
        if hit is not None and hit[0] == sig and all(r() is p for r, p in zip(hit[2], params)):
            if _CAPTURE_REFS is not None:
                _CAPTURE_REFS.append(hit[1])
            return hit[1]
        if hit is not None:
            bump_generation()
        with torch.no_grad():
            val = fn(*params)
        if _CAPTURE_REFS is not None:
            _DERIVED_IN_CAPTURE[0] = True
            _CAPTURE_REFS.append(val)
            return val
        if len(self._d) > 4096:                       # drop entries whose parameters are gone
            self._d = {kk: v for kk, v in self._d.items() if all(r() is not None for r in v[2])}
        self._d[k] = (sig, val, tuple(weakref.ref(p) for p in params), None)
        return val


CACHE = ParamCache()


def w_c(p):
    """weight as a contiguous 2-D [N, K] matrix in the compute dtype."""
    dt = compute_dtype()
    return CACHE.get(p, ("wc", dt), lambda t: t.detach().reshape(t.shape[0], -1).to(dt).contiguous())


def p32(p):
    """parameter as contiguous fp32 (biases, LayerNorm affine, tables)."""
    return CACHE.get(p, "p32", lambda t: t.detach().float().contiguous())


def conv_w_c(p):
    """conv weight [Cout,Cin,k,k] -> [Cout, k*k*Cin] (tap order ky,kx,ci) in the compute dtype."""
    dt = compute_dtype()
    return CACHE.get(p, ("convw", dt),
                     lambda t: t.detach().permute(0, 2, 3, 1).reshape(t.shape[0], -1).to(dt).contiguous())


def convT_w_c(p):
    """ConvTranspose2d k2s2 weight [Cin,Cout,2,2] -> GEMM weight [(dy,dx,co), Cin] in the compute dtype."""
    dt = compute_dtype()
    return CACHE.get(p, ("convTw", dt),
                     lambda t: t.detach().permute(2, 3, 1, 0).reshape(-1, t.shape[0]).to(dt).contiguous())


def cat_w_c(*ps):
    dt = compute_dtype()
    return CACHE.get(ps, ("catw", dt), lambda *ts: torch.cat([t.detach().reshape(t.shape[0], -1) for t in ts], 0).to(dt).contiguous())


def cat_p32(*ps):
    return CACHE.get(ps, "cat32", lambda *ts: torch.cat([t.detach().float().reshape(-1) for t in ts], 0).contiguous())


# ---------------------------------------------------------------- parallel branches inside a captured graph
_CAPTURING = False
_IN_PAR = False
_BRANCH_STREAMS = []


def par(*fns):
    """Run independent closures and return their results.  Eagerly they simply run one after the other; while a
    GraphRunner is capturing, every closure after the first runs on its own stream that forks from / joins the
    capturing stream, so the captured graph has parallel branches (the SAM heads are ~115 dependent 2-12 us kernels:
    image-side and token-side projections, the up-scaling path and the head MLPs do not depend on each other)."""
    global _IN_PAR
    if not _CAPTURING or _IN_PAR or len(fns) < 2:
        return [f() for f in fns]
    main = torch.cuda.current_stream()
    while len(_BRANCH_STREAMS) < len(fns) - 1:
        _BRANCH_STREAMS.append(torch.cuda.Stream())
    outs = [None] * len(fns)
    _IN_PAR = True
    try:
        for i, f in enumerate(fns[1:]):
            s = _BRANCH_STREAMS[i]
            s.wait_stream(main)
            with torch.cuda.stream(s):
                outs[i + 1] = f()
        outs[0] = fns[0]()
        for i in range(len(fns) - 1):
            main.wait_stream(_BRANCH_STREAMS[i])
    finally:
        _IN_PAR = False
    return outs


class GraphRunner:
    """CUDA-graph replay of a fixed-shape sub-pipeline of native kernels (the mask decoder is
    ~170 tiny launches per slice and launch-bound).  `run(key, fn, tensors, clone=True)` executes
    `fn(*tensors)`: the first `warmup` calls per key run eagerly (parameter caches, tables and kernel
    attributes get initialised outside any capture), then the call is captured once on static copies of the
    inputs and replayed afterwards.  Inputs keep their strides (NCHW-shaped views of NHWC memory stay views);
    outputs live in the graph's private pool and are cloned out unless the caller consumes them immediately.
    Shapes/strides/dtypes, `key` and the compute dtype select a graph - values never do, so the kernels must not read
    host state that changes between calls.  A graph is only replayed while the parameter generation it was captured
    under is current (`bump_generation`: load_state_dict, train()/eval() switches, `.to()`, a re-derived parameter
    copy); otherwise it is dropped and re-captured after a fresh eager warm-up.  The graph entry keeps strong references
    to every derived parameter copy its capture touched."""

    def __init__(self, warmup=2):
        self.warmup = warmup
        self._seen = {}
        self._graphs = {}
        self.replays = 0

    @staticmethod
    def _sig(t):
        return None if t is None else (tuple(t.shape), tuple(t.stride()), t.dtype)

    def run(self, key, fn, tensors, clone=True):
        from . import native, ops
        full_key = (key, compute_dtype(), tuple(self._sig(t) for t in tensors))
        entry = self._graphs.get(full_key)
        if entry is not None and entry[5] != _GEN[0]:
            del self._graphs[full_key]                 # parameters changed since the capture: never replay stale weights
            entry = None
        if entry is None:
            # eager warm-up calls are counted PER parameter generation: the warm-up is what re-derives the kernel-ready
            # parameter copies outside of any capture
            g, n = self._seen.get(full_key, (_GEN[0], 0))
            if g != _GEN[0]:
                n = 0
            self._seen[full_key] = (_GEN[0], n + 1)
            if n < self.warmup or not torch.cuda.is_available():
                return fn(*tensors)
            # empty_like keeps the strides of dense (permuted) views and densifies expanded (stride-0) ones
            static_in = [None if t is None else torch.empty_like(t) for t in tensors]
            for s, t in zip(static_in, tensors):
                if s is not None:
                    s.copy_(t)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            n0 = native.launch_count
            global _CAPTURING, _CAPTURE_REFS
            while len(_BRANCH_STREAMS) < 2:            # branch streams of par() exist before the capture starts
                _BRANCH_STREAMS.append(torch.cuda.Stream())
            _CAPTURING = True
            _CAPTURE_REFS = refs = []
            _DERIVED_IN_CAPTURE[0] = False
            gen0 = _GEN[0]
            try:
                with torch.cuda.graph(graph):
                    static_out = fn(*static_in)
            finally:
                _CAPTURING = False
                _CAPTURE_REFS = None
            if _GEN[0] != gen0 or _DERIVED_IN_CAPTURE[0]:
                # a parameter copy was (re-)derived DURING the capture: throw the capture away, warm up again
                self._seen[full_key] = (_GEN[0], 0)
                del graph
                return fn(*tensors)
            entry = self._graphs[full_key] = [graph, static_in, static_out, native.launch_count - n0, None, gen0, refs]
        graph, static_in, static_out, launches, last = entry[:5]
        cur = torch.cuda.current_stream()
        if last is not None and last[0] != cur.cuda_stream:
            # the previous replay ran on another stream (slice encoding ahead of need): the static buffers are shared
            cur.wait_event(last[1])
        # all inputs move into the static buffers in one launch (they keep the callers' strides: byte copies)
        ops.multi_copy([(s, t) for s, t in zip(static_in, tensors) if s is not None])
        graph.replay()
        self.replays += 1
        native.launch_count += launches            # kernels of this library executed by the replay
        out = static_out if not clone else _clone_tree(static_out)
        ev = last[1] if last is not None else torch.cuda.Event()
        ev.record(cur)
        entry[4] = (cur.cuda_stream, ev)
        return out


def _clone_tree(x, memo=None, pairs=None):
    """clone every tensor of a nested result ONCE (the SAM heads return the same mask tensor under two names when a
    single mask is requested: two 4 MB copies per slice otherwise); all clones are filled by one ms2_multi_copy launch."""
    from . import ops
    top = memo is None
    if top:
        memo, pairs = {}, []
    if isinstance(x, torch.Tensor):
        k = id(x)
        if k not in memo:
            if x.is_cuda and ops._dense(x):
                memo[k] = torch.empty_like(x)          # preserve_format: same strides for dense tensors
                pairs.append((memo[k], x))
            else:
                memo[k] = x.clone()
        out = memo[k]
    elif isinstance(x, (list, tuple)):
        out = type(x)(_clone_tree(v, memo, pairs) for v in x)
    elif isinstance(x, dict):
        out = {k: _clone_tree(v, memo, pairs) for k, v in x.items()}
    else:
        out = x
    if top and pairs:
        ops.multi_copy(pairs)
    return out

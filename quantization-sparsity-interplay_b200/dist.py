"""Multi-GPU partitioning of the hot path.  One process per GPU (torchrun); torch.distributed is
plumbing only.

* The whole-model compression pass shards BY TENSOR: quantiser units are independent per block, so ranks never exchange
  data -- no collective on the data path.  Layer l of the model goes to rank l % world_size.
* Timing is the max over ranks of the device time (all_reduce MAX), never wall clock.
Works with backend 'nccl' (GPU boxes) and 'gloo' (CPU tests of this host logic).
"""
import os

import torch
import torch.distributed as dist

# (N_out, K) of the BFPLinear weights of one decoder layer
LAYER_SHAPES = {
    "llama-7b": [(4096, 4096)] * 4 + [(11008, 4096)] * 2 + [(4096, 11008)],
    "llama-13b": [(5120, 5120)] * 4 + [(13824, 5120)] * 2 + [(5120, 13824)],
    "llama-65b": [(8192, 8192)] * 4 + [(22016, 8192)] * 2 + [(8192, 22016)],
    "opt-66b": [(9216, 9216)] * 4 + [(36864, 9216), (9216, 36864)],
}
NUM_LAYERS = {"llama-7b": 32, "llama-13b": 40, "llama-65b": 80, "opt-66b": 64}


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init(backend=None):
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"), rank=rank, world_size=world)
    return rank, local_rank, world


def model_tensors(model, num_layers=None):
    """[(layer, index_in_layer, (N_out, K))] of every BFPLinear weight of `model`."""
    L = NUM_LAYERS[model] if num_layers is None else num_layers
    return [(l, i, s) for l in range(L) for i, s in enumerate(LAYER_SHAPES[model])]


def shard_by_layer(tensors, rank, world):
    """The tensors rank `rank` compresses: layer l -> rank l % world.  A partition: disjoint, covering, no exchange."""
    return [t for t in tensors if t[0] % world == rank]


def shutdown():
    """Tear the process group down (quietens NCCL's leak warning at exit); a no-op when none was created."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def barrier(device=None):
    if dist.is_initialized():
        if device is not None and device.type == "cuda":
            dist.barrier(device_ids=[device.index])
        else:
            dist.barrier()


def max_over_ranks(x, device=None):
    """max over ranks of a python float (device time in ms)."""
    if not dist.is_initialized():
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x, device=None):
    if not dist.is_initialized():
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


# ---------------------------------------------------------------------------------------------------------------
# Column-parallel BFP linear: W[N_out, K] is split by rows of N_out into contiguous shards.
# Blocks and N:M groups run along K, so a shard quantises exactly as the same rows of the full matrix do; x is
# replicated and quantised redundantly; rank g computes y_g = x W_g^T [T, N_out/G]; ONE all-gather assembles y.
# ---------------------------------------------------------------------------------------------------------------
def column_shard(n_out, rank, world):
    """[lo, hi) rows of N_out owned by `rank` (contiguous, sizes differ by at most one)."""
    base, rem = divmod(n_out, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def bind_to_gpu_numa(local_rank):
    """Pins this process (and therefore the first-touch placement of the pinned host buffers it allocates afterwards) to the
    CPUs NVML reports as local to GPU `local_rank`.  With one process per GPU this keeps every rank's H2D / D2H staging on its
    own socket instead of all ranks sharing NUMA node 0.  Returns the CPU list, or None when NVML / affinity is unavailable
    (single-socket hosts, containers without NVML): the caller then simply keeps the inherited affinity."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        idx = local_rank
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            if local_rank < len(ids) and ids[local_rank].isdigit():
                idx = int(ids[local_rank])
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


class _SymmetricOutputs:
    """Two symmetric (torch.distributed._symmetric_memory) [capacity, N] output buffers per (dtype, device), alternating between
    calls, allocated for the largest token count seen so far (growth re-allocates on EVERY rank at the same call, because all
    ranks see the same T) and sliced for smaller T -- a new sequence length no longer re-rendezvouses on the hot path."""

    def __init__(self, group, n_out):
        self.group, self.n_out, self.slots = group, n_out, {}

    def get(self, T, dtype, device):
        import torch.distributed._symmetric_memory as symm_mem
        key = (dtype, device)
        slot = self.slots.get(key)
        if slot is None or slot["cap"] < T:
            cap = max(T, 2 * slot["cap"] if slot else 0)
            bufs = []
            for _ in range(2):
                b = symm_mem.empty((cap, self.n_out), dtype=dtype, device=device)
                bufs.append((b, symm_mem.rendezvous(b, self.group)))
            slot = self.slots[key] = {"cap": cap, "bufs": bufs, "turn": 0}
        buf, hdl = slot["bufs"][slot["turn"]]
        slot["turn"] ^= 1
        return buf, hdl


class ColumnParallelBFPLinear(torch.nn.Module):
    """Drop-in for BFPLinear(in_features, out_features) on `world` GPUs: each rank holds rows column_shard(out_features)
    of the weight (and bias) in a local BFPLinear and the forward all-gathers the output slices.

    Which gather runs is decided ONCE and COLLECTIVELY (first CUDA forward, or `decide_path()`): every rank probes whether the
    fused path (all-gather in the GEMM epilogue over peer memory) is available to it -- configuration, alignment, symmetric
    memory allocation -- and the flags are MIN-reduced, so either all ranks take the fused path or all take NCCL.  After that
    nothing is caught inside forward: a launch error on one rank propagates instead of silently desynchronising the
    collectives of the others."""

    def __init__(self, in_features, out_features, bias=True, group=None, **bfp_kwargs):
        super().__init__()
        from .bfp_ops import BFPLinear
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.in_features, self.out_features = in_features, out_features
        self.lo, self.hi = column_shard(out_features, self.rank, self.world)
        self.local = BFPLinear(in_features, self.hi - self.lo, bias=bias, **bfp_kwargs)
        self._symm = None            # _SymmetricOutputs of the fused path
        self._path = None            # None = undecided, "fused" or "nccl" (same on every rank)
        self._fused_failed = None    # why the fused path was ruled out (for reports)

    @torch.no_grad()
    def load_full(self, weight, bias=None):
        """Copies this rank's shard out of the full [N_out, K] weight (and [N_out] bias)."""
        self.local.weight.copy_(weight[self.lo:self.hi])
        if bias is not None and self.local.bias is not None:
            self.local.bias.copy_(bias[self.lo:self.hi])
        self.local.invalidate_packed()
        return self

    # ---- fused path: the all-gather happens in the GEMM epilogue (peer-memory TMA stores), no NCCL call ----------------
    def _fused_static_ok(self, x):
        """Conditions that are identical on every rank by construction: 2:4 sparse tensor-core kind, nearest rounding, CUDA,
        more than one rank, every slice 16-byte aligned."""
        from . import bfp_ops
        if self.world == 1 or not x.is_cuda or os.environ.get("BFP_COLUMN_PARALLEL", "fused") != "fused":
            return "disabled (single rank, CPU tensor or BFP_COLUMN_PARALLEL != fused)"
        if self.local.bfp_args['rounding_mode'] != bfp_ops.rounding_modes.DETERM or self.local.num_format != 'bfp':
            return "needs the bfp format with nearest rounding"
        if bfp_ops._tensor_core_kind(x, self.local.weight, self.local.bfp_args) != 'sp':
            return "weight is not on the 2:4 sparse tensor-core kind"
        if self.out_features % 8 or any(column_shard(self.out_features, r, self.world)[0] % 8 for r in range(self.world)):
            return "output slices are not 16-byte aligned"
        return None

    def decide_path(self, x):
        """Collective (call on every rank with the same shapes): picks "fused" or "nccl" for this module, once."""
        if self._path is not None:
            return self._path
        why = self._fused_static_ok(x)
        ok = why is None
        if ok:
            try:                                                    # per-rank probe: symmetric memory may be missing on one rank only
                group = self.group if self.group is not None else dist.group.WORLD
                self._symm = _SymmetricOutputs(group, self.out_features)
            except Exception as e:                                  # noqa: BLE001  (import-time failure only; allocation is below)
                ok, why = False, repr(e)
        if self.world > 1 and x.is_cuda and os.environ.get("BFP_COLUMN_PARALLEL", "fused") == "fused":
            # every rank reaches this all_reduce whatever its local answer was
            probe_ok = ok
            if ok:
                try:
                    K = x.shape[-1]
                    self._symm.get(max(1, x.numel() // K), x.dtype, x.device)
                except Exception as e:                              # noqa: BLE001
                    probe_ok, why = False, "symmetric memory unavailable: " + repr(e)[:200]
            flag = torch.tensor([1 if probe_ok else 0], dtype=torch.int32, device=x.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            ok = bool(flag.item())
            if not ok and why is None:
                why = "another rank cannot take the fused path"
        self._path = "fused" if ok else "nccl"
        self._fused_failed = None if ok else why
        if not ok:
            self._symm = None
        return self._path

    def _fused_launch(self, x, buf, hdl):
        """Enqueues this module's GEMM whose epilogue stores the tile into ALL ranks' full outputs."""
        import ctypes
        from . import _lib, bfp_ops
        K = x.shape[-1]
        T = x.numel() // K
        xb = bfp_ops._packed_activation(x, self.local.bfp_args)
        ws = self.local._packed_weight('sp')
        bias = self.local.bias.detach().float().contiguous() if self.local.bias is not None else None
        es = buf.element_size()
        ptrs = (ctypes.c_void_p * self.world)(*[int(hdl.buffer_ptrs[r]) + self.lo * es for r in range(self.world)])
        with torch.cuda.device(x.device):
            _lib.check(_lib.lib().bfp_gemm_bf16_sp_gather(xb.data_ptr(), ws.comp.data_ptr(), ws.meta.data_ptr(),
                                                          bias.data_ptr() if bias is not None else None, ptrs, self.world, bfp_ops._DT[x.dtype],
                                                          self.out_features, T, self.hi - self.lo, xb.shape[1], torch.cuda.current_stream().cuda_stream))

    def _fused_forward(self, x, alias_output=False):
        """y = [x W_0^T | ... | x W_{G-1}^T]: every rank runs bfp_gemm_bf16_sp_gather on its shard and the epilogue stores the
        tile into ALL ranks' full outputs (torch symmetric memory: own HBM + peers over NVLink).  Two symmetric buffers
        alternate so a barrier before the kernel is enough to know the target is free; a barrier after it makes every
        slice visible everywhere.  alias_output=True returns the symmetric buffer itself (valid until the second-next call)."""
        K = x.shape[-1]
        T = x.numel() // K
        buf, hdl = self._symm.get(T, x.dtype, x.device)
        hdl.barrier(channel=0)                                       # nobody is still reading this buffer's previous contents
        self._fused_launch(x, buf, hdl)
        hdl.barrier(channel=1)                                       # every rank's slices have landed in every buffer
        out = buf[:T] if alias_output else buf[:T].clone()
        return out.view(tuple(x.shape[:-1]) + (self.out_features,))

    def forward(self, x, alias_output=False):
        inference = not (torch.is_grad_enabled() and (x.requires_grad or self.local.weight.requires_grad))
        if self.world > 1 and x.is_cuda and inference and self.decide_path(x) == "fused":
            return self._fused_forward(x, alias_output)             # errors propagate: the other ranks are in the same path
        y = self.local(x)                                           # [..., N_local]
        if self.world == 1:
            return y
        lead = y.shape[:-1]
        y2 = y.reshape(-1, y.shape[-1]).contiguous()
        sizes = [column_shard(self.out_features, r, self.world) for r in range(self.world)]
        wmax = max(hi - lo for lo, hi in sizes)
        equal = all(hi - lo == wmax for lo, hi in sizes)
        if not equal:                                               # collectives want equal shapes: pad the narrow shards
            y2 = torch.nn.functional.pad(y2, (0, wmax - y2.shape[1]))
        buf = torch.empty((self.world,) + tuple(y2.shape), dtype=y2.dtype, device=y2.device)
        if y2.is_cuda:
            dist.all_gather_into_tensor(buf, y2, group=self.group)  # NCCL over NVLink: [G, T, N/G]
        else:
            dist.all_gather(list(buf.unbind(0)), y2, group=self.group)
        if equal:
            out = buf.permute(1, 0, 2).reshape(y2.shape[0], self.out_features)
        else:
            out = torch.cat([buf[r, :, : hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=1)
        return out.reshape(lead + (self.out_features,))


def column_parallel_group_forward(mods, x):
    """Sibling column-parallel projections of the SAME input (q/k/v, gate/up) as one unit: ONE barrier, the fused GEMMs
    enqueued back to back, ONE barrier.  The peer stores of projection i (NVLink) then run under the GEMM of projection i+1
    instead of every projection paying its own exposed gather and two barriers.  All modules must have decided "fused"
    (collectively, `decide_path`); otherwise each is run on its own.  Returns the outputs as views of the symmetric
    buffers (valid until each module's second-next call), bit-identical to calling the modules one by one."""
    if not mods:
        return []
    inference = not (torch.is_grad_enabled() and (x.requires_grad or any(m.local.weight.requires_grad for m in mods)))
    if not (x.is_cuda and inference and all(m.world > 1 and m.decide_path(x) == "fused" for m in mods)):
        return [m(x) for m in mods]
    K = x.shape[-1]
    T = x.numel() // K
    slots = [m._symm.get(T, x.dtype, x.device) for m in mods]
    slots[0][1].barrier(channel=0)                                   # one group-wide barrier covers every module's buffer
    for m, (buf, hdl) in zip(mods, slots):
        m._fused_launch(x, buf, hdl)
    slots[0][1].barrier(channel=1)
    return [buf[:T].view(tuple(x.shape[:-1]) + (m.out_features,)) for m, (buf, _) in zip(mods, slots)]

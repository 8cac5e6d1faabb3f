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

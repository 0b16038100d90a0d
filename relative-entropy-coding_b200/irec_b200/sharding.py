"""Host-side sharding logic of the multi-GPU paths (no device code here; backend-agnostic, so the
world_size > 1 plumbing is tested on CPU with gloo -- tests/test_sharding_gloo.py).

Two ways the iREC path shards:
  * independent units: images / coder-blocks are split into contiguous ranges, no collective in the loop;
  * one exchange per auxiliary variable: the candidate index range [0, S) of ONE coder-block is split
    contiguously, every rank keeps its local top-B records and the records are all-gathered.

Record layout (matches irec_record_t of include/irec.h): 4 x int32 = (float32 score bits, s, b, pad).
The per-rank exchange buffer is (B + 1) records: B records followed by one record whose first word is the
number of valid records (so a single all_gather moves everything).
"""
import torch

RECORD_WORDS = 4


def unit_range(n_units, rank, world):
    """contiguous range of independent units (images, blocks) of `rank`; the remainder is spread over the first ranks"""
    base, rem = divmod(int(n_units), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def candidate_range(S, rank, world):
    """contiguous candidate-sample range [s_begin, s_end) of `rank` (ceil split; trailing ranks may be empty)"""
    per = (int(S) + world - 1) // world
    lo = min(int(S), rank * per)
    return lo, min(int(S), lo + per)


def exchange_words(B):
    return (int(B) + 1) * RECORD_WORDS


def alloc_exchange(B, world, device):
    """(local, gathered) int32 buffers of the record exchange"""
    local = torch.zeros(exchange_words(B), dtype=torch.int32, device=device)
    gathered = torch.zeros(world * exchange_words(B), dtype=torch.int32, device=device)
    return local, gathered


def exchange_records(local, gathered, B, world, dist=None, group=None):
    """all-gather of the per-rank (records, count) buffers -> (records [world, B*4] int32, counts [world] int32).
    With world == 1 no collective is issued."""
    w = exchange_words(B)
    if world > 1:
        dist.all_gather_into_tensor(gathered, local, group=group)
        g = gathered.view(world, w)
    else:
        g = local.view(1, w)
    records = g[:, :B * RECORD_WORDS].contiguous()
    counts = g[:, B * RECORD_WORDS].contiguous()
    return records, counts

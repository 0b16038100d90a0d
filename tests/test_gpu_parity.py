"""GPU parity tests (run on a B200 with -m gpu): every CUDA path, through the C ABI, against the CPU oracle
on the same seeded inputs.  Bar: bit-exact indices, bit-exact samples, bit-exact decode."""
import os

import numpy as np
import pytest

import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda(built):
    import torch
    assert torch.cuda.is_available(), "these tests need a GPU"
    from irec_b200 import native
    native.lib()
    return torch.device("cuda:0")


def bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


def to_dev(arrs, dev):
    import torch
    return [torch.as_tensor(a, dtype=torch.float32, device=dev).contiguous() for a in arrs]


# ------------------------------------------------------------------------------------------ tables / streams
def test_ndtri_table_bit_exact(cuda):
    from irec_b200 import native
    T = native.ndtri_table().numpy()
    assert np.array_equal(bits(T[1:]), bits(O.ndtri_table()[1:]))


@pytest.mark.parametrize("q,start,n", [(42, 0, 4096), (43, 5, 1001), (69420, 123457, 3000), (0, 0, 64),
                                        (2 ** 31 - 1, 0, 64), (2 ** 33 + 7, 3, 64)])
def test_uniform_int_stream(cuda, q, start, n):
    from irec_b200 import engine
    r = engine.beam_uniform_int(q, start, n).cpu().numpy()
    assert np.array_equal(r, O.beam_uniform_int(q, start, n))


@pytest.mark.parametrize("seed,start,n", [(42, 0, 8192), (69420, 7, 1001), (0, 0, 256)])
def test_normal_stream(cuda, seed, start, n):
    from irec_b200 import engine
    z = engine.is_normal_stream(seed, start, n).cpu().numpy()
    assert np.array_equal(bits(z), bits(O.is_normal_stream(seed, start, n)))


def test_kl_naux(cuda):
    import torch
    from irec_b200 import engine
    tl, ts, pl, ps = synth.c2(8192, data_seed=3)
    d = to_dev((tl, ts, pl, ps), cuda)
    offs, nb, _ = engine.make_block_offsets(8192, 1000, cuda)
    kl, na = engine.kl_naux(*d, None, offs, nb, 3.0)
    kl, na = kl.cpu().numpy(), na.cpu().numpy()
    for b in range(nb):
        lo, hi = 1000 * b, min(8192, 1000 * b + 1000)
        ok = O.kl(tl[lo:hi], ts[lo:hi], pl[lo:hi], ps[lo:hi])
        assert bits(kl[b]) == bits(np.float32(ok))
        assert na[b] == O.n_aux(ok, 3.0)


# ------------------------------------------------------------------------------------------ beam coder
BEAM_CASES = [
    # recipe, D, data_seed, omega, extra, B, seed
    ("c1", 64, 0, 3.0, 1.2, 1, 42),          # BASELINE config 1
    ("c1", 64, 0, 3.0, 1.2, 20, 42),
    ("c1", 64, 1, 3.0, 1.0, 10, 69420),
    ("c2", 1000, 1, 3.0, 1.2, 20, 42),       # one C2 coder-block
    ("c2", 192, 2, 3.0, 1.2, 20, 42),        # last C2 block
    ("c3", 1000, 4, 3.0, 1.0, 10, 7),        # one C3 coder-block
    ("c3", 288, 3, 3.0, 1.0, 10, 7),
    ("c3", 56, 5, 3.0, 1.0, 10, 7),
    ("c2", 37, 6, 3.0, 1.2, 5, 1),           # D not a multiple of 4: unaligned Philox quads
    ("c2", 1, 7, 1.0, 1.0, 3, 1),
    ("c2", 1023, 8, 4.0, 1.0, 2, 11),
    ("c2", 257, 9, 2.0, 1.0, 32, 5),         # S=7 < B: the beam count grows 1 -> 7 -> 32
    ("c2", 100, 10, 5.0, 1.3, 16, 2 ** 31 + 5),
]


def run_beam_case(cuda, recipe, D, data_seed, omega, extra, B, seed):
    import torch
    from irec_b200 import Normal
    from rec.coding import BeamSearchCoder
    tl, ts, pl, ps = getattr(synth, recipe)(D, data_seed=data_seed)
    coder = BeamSearchCoder(kl_per_partition=omega, n_beams=B, extra_samples=extra)
    ref = O.beam_encode_block(tl, ts, pl, ps, omega, coder.n_samples, B, seed)
    t = Normal(tl[None, :], ts[None, :], device=cuda)
    p = Normal(pl[None, :], ps[None, :], device=cuda)
    indices, sample = coder.encode(t, p, seed=seed)
    assert list(indices) == ref["indices"].tolist()
    assert np.array_equal(bits(sample.cpu().numpy().reshape(-1)), bits(ref["sample"]))
    assert coder.get_codelength(indices) == len(ref["indices"]) * np.log(coder.n_samples)
    keep = list(indices)
    dec = coder.decode(p, indices, seed=seed)
    assert indices == keep[::-1], "decode_block reverses the caller's list in place (reference :127)"
    assert torch.equal(dec, sample)
    odec = O.beam_decode_block(pl, ps, coder.n_samples, seed, ref["indices"])
    assert np.array_equal(bits(dec.cpu().numpy().reshape(-1)), bits(odec))


KERNEL_ENV = {
    "resident1": {"IREC_RESIDENT": "1"},                              # k_beam_encode_resident
    "resident2": {"IREC_RESIDENT": "2"},                              # k_beam_encode_resident2 (discrete-log table addressing)
    "resident2-notable": {"IREC_RESIDENT": "2", "IREC_R2_NO_TABLE": "1"},   # exponents from Philox + dlog in place
    "tmem": {"IREC_RESIDENT": "3"},                                   # k_beam_encode_tmem (beams in tensor memory, two contexts per SM)
    "tmem-notable": {"IREC_RESIDENT": "3", "IREC_R2_NO_TABLE": "1"},
    "cluster8": {"IREC_CLUSTER": "8"},                                # k_beam_encode_cluster, 8 CTAs per coder-block
    "cluster4": {"IREC_CLUSTER": "4"},
    "cluster8-notable": {"IREC_CLUSTER": "8", "IREC_R2_NO_TABLE": "1"},
}


class kernel_env:
    def __init__(self, kernel):
        self.env = KERNEL_ENV[kernel]

    def __enter__(self):
        for k in ("IREC_FORCE_GENERAL", "IREC_RESIDENT", "IREC_R2_NO_TABLE", "IREC_CLUSTER"):
            os.environ.pop(k, None)
        os.environ.update(self.env)

    def __exit__(self, *exc):
        for k in self.env:
            os.environ.pop(k, None)


@pytest.mark.parametrize("kernel", list(KERNEL_ENV))
@pytest.mark.parametrize("case", BEAM_CASES, ids=[f"{c[0]}-D{c[1]}-B{c[5]}" for c in BEAM_CASES])
def test_beam_resident_vs_oracle(cuda, case, kernel):
    """both generations of the persistent per-block kernel and the cluster-per-block kernel (with and without the
    per-launch exponent table) are bit-identical to the oracle"""
    with kernel_env(kernel):
        if kernel.startswith("tmem"):
            from irec_b200 import native as N
            S = int(np.exp(case[3] * case[4]))
            if N.lib().irec_beam_encode_path(1, case[1], S, case[5]) != 3:
                pytest.skip("sizes not covered by k_beam_encode_tmem (2 <= n_beams <= 20, S * 20 scores per context): resident2 runs them")
        run_beam_case(cuda, *case)


WIDE_CASES = [
    # recipe, D, data_seed, omega, extra, B, seed          n_beams > 32: k_beam_encode_wide (csrc/irec_wide.cu)
    ("c1", 64, 0, 3.0, 1.2, 33, 42),         # one more than the persistent kernels take
    ("c2", 192, 2, 3.0, 1.2, 64, 42),
    ("c2", 1000, 1, 3.0, 1.2, 100, 42),      # a full C2 coder-block, B not a multiple of the scoring page
    ("c2", 257, 9, 2.0, 1.0, 300, 5),        # S = 7: the beam count grows 1 -> 7 -> 49 -> 300
    ("c3", 288, 3, 3.0, 1.0, 1024, 7),       # the largest beam width (S = 20: 1 -> 20 -> 400 -> 1024)
    ("c2", 37, 6, 4.0, 1.3, 40, 1),          # unaligned Philox quads
]


@pytest.mark.parametrize("case", WIDE_CASES, ids=[f"{c[0]}-D{c[1]}-B{c[5]}" for c in WIDE_CASES])
def test_beam_wide_vs_oracle(cuda, case):
    """n_beams > 32 (the reference has no limit, beam_search_coder.py:15-30): beams, scores and history in global memory,
    radix-select top-B -- indices, sample bits and decode identical to the oracle"""
    from irec_b200 import native as N
    S = int(np.exp(case[3] * case[4]))
    assert N.lib().irec_beam_encode_path(1, case[1], S, case[5]) == 4
    run_beam_case(cuda, *case)


def test_beam_wide_batch_ragged(cuda):
    """several coder-blocks of different sizes in one wide-beam launch (persistent CTAs over the block queue)"""
    import torch
    from irec_b200 import engine
    sizes = [100, 200, 37, 1000, 200, 5]
    B, omega, S, seed = 48, 3.0, 36, 42
    arrs = [synth.c2(n, data_seed=40 + i) for i, n in enumerate(sizes)]
    flat = [torch.from_numpy(np.concatenate([a[k] for a in arrs])).to(cuda) for k in range(4)]
    offsets = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int64, device=cuda)
    out = engine.beam_encode_blocks(*flat, None, offsets, len(sizes), max(sizes), omega, S, B, seed)
    smp = out.sample.cpu().numpy()
    lo = 0
    for b, (n, a) in enumerate(zip(sizes, arrs)):
        ref = O.beam_encode_block(*a, omega, S, B, seed)
        assert out.indices[b] == ref["indices"].tolist()
        assert np.array_equal(bits(smp[lo:lo + n]), bits(ref["sample"]))
        lo += n


def test_beam_too_wide_is_an_error(cuda):
    """beyond the wide kernel's range the call fails loudly (no fallback)"""
    from irec_b200 import native as N
    assert N.lib().irec_beam_encode_path(1, 64, 36, 1025) == 0
    assert N.lib().irec_beam_encode_path(1, 2000, 36, 64) == 0


@pytest.mark.parametrize("kernel", ["resident2", "tmem", "cluster8", "cluster4"])
def test_beam_ragged_blocks_one_launch(cuda, kernel):
    """four different block sizes in ONE launch: two get a per-launch exponent table, the others generate their
    candidate exponents in place; every block must equal the oracle on its own slice"""
    with kernel_env(kernel):
        _ragged_blocks(cuda)


def test_default_kernel_choice(cuda):
    """few coder-blocks per launch (one image: 9 / 13 blocks) -> the cluster kernel; a batch -> the persistent kernel"""
    from irec_b200 import native as N
    lib = N.lib()
    for k in ("IREC_FORCE_GENERAL", "IREC_RESIDENT", "IREC_CLUSTER"):
        os.environ.pop(k, None)
    assert lib.irec_beam_encode_path(9, 1000, 36, 20) == 108        # 100 + cluster size
    assert lib.irec_beam_encode_path(13, 1000, 20, 10) == 108
    assert lib.irec_beam_encode_path(24, 1000, 36, 20) == 104
    assert lib.irec_beam_encode_path(302, 1000, 20, 10) == 3        # tensor-memory kernel
    assert lib.irec_beam_encode_path(1152, 1000, 36, 20) == 3
    assert lib.irec_beam_encode_path(1152, 1000, 36, 32) == 1       # n_beams > 20: the first-generation kernel (a 32-beam matrix does not fit beside the 120 KB table)
    assert lib.irec_beam_encode_path(1152, 256, 36, 32) == 2        # ... resident2 where it does
    assert lib.irec_beam_encode_path(1, 64, 36, 1) == 2
    assert lib.irec_beam_encode_path(1, 2500, 36, 4) == 0


def _ragged_blocks(cuda):
    import torch
    from irec_b200 import engine
    sizes = [100, 200, 37, 1000, 200, 100, 5]
    n = sum(sizes)
    tl, ts, pl, ps = synth.c2(n, data_seed=77)
    d = to_dev((tl, ts, pl, ps), cuda)
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    offsets = torch.as_tensor(offs, device=cuda)
    res = engine.beam_encode_blocks(*d, None, offsets, len(sizes), max(sizes), 3.0, 36, 20, 42)
    out = res.sample.cpu().numpy()
    for b, D in enumerate(sizes):
        lo, hi = offs[b], offs[b + 1]
        ref = O.beam_encode_block(tl[lo:hi], ts[lo:hi], pl[lo:hi], ps[lo:hi], 3.0, 36, 20, 42)
        assert res.indices[b] == ref["indices"].tolist(), (b, D)
        assert np.array_equal(bits(out[lo:hi]), bits(ref["sample"])), (b, D)


@pytest.mark.parametrize("case", BEAM_CASES[:9], ids=[f"{c[0]}-D{c[1]}-B{c[5]}" for c in BEAM_CASES[:9]])
def test_beam_general_path_vs_oracle(cuda, case):
    os.environ["IREC_FORCE_GENERAL"] = "1"
    try:
        run_beam_case(cuda, *case)
    finally:
        os.environ.pop("IREC_FORCE_GENERAL", None)


def test_beam_large_dim_general_path(cuda):
    """D > 1024 (no block_size): multi-slot reduction tree"""
    run_beam_case(cuda, "c2", 2500, 12, 6.0, 1.0, 4, 3)


def test_reference_unit_test_case(cuda):
    """rec/coding/tests/test_coder.py:12-21"""
    import torch
    from irec_b200 import Normal
    from rec.coding import BeamSearchCoder
    encoder = BeamSearchCoder(kl_per_partition=6., n_beams=10, extra_samples=1.)
    t = Normal(torch.tensor([[5.1]]), torch.tensor([[0.001]]), device=cuda)
    p = Normal(torch.tensor([[0.]]), torch.tensor([[1.]]), device=cuda)
    indices, sample = encoder.encode(t, p, seed=69420, update_sampler=False)
    ref = O.beam_encode_block([5.1], [0.001], [0.], [1.], 6., 403, 10, 69420)
    assert list(indices) == ref["indices"].tolist()
    reconstructed = encoder.decode(p, indices, seed=69420)
    assert torch.equal(sample, reconstructed)
    assert abs(float(sample) - 5.1) < 0.01


def test_beam_block_size_split(cuda):
    """whole-tensor encode with block_size: Coder.split permutation + all blocks in one launch (C2 level shape)"""
    import torch
    from irec_b200 import Normal
    from rec.coding import BeamSearchCoder
    n, bs, seed = 8192, 1000, 42
    tl, ts, pl, ps = synth.c2(n, data_seed=21)
    coder = BeamSearchCoder(kl_per_partition=3., n_beams=20, extra_samples=1.2, block_size=bs)
    shape = (1, 16, 16, 32)
    t = Normal(tl.reshape(shape), ts.reshape(shape), device=cuda)
    p = Normal(pl.reshape(shape), ps.reshape(shape), device=cuda)
    indices, sample = coder.encode(t, p, seed=seed)
    assert sample.shape == shape and len(indices) == 9
    perm = O.shuffle_perm(n, seed)
    out = np.zeros(n, np.float32)
    for b in range(9):
        sel = perm[b * bs:min(n, (b + 1) * bs)]
        ref = O.beam_encode_block(tl[sel], ts[sel], pl[sel], ps[sel], 3., 36, 20, seed)
        assert indices[b] == ref["indices"].tolist(), b
        out[sel] = ref["sample"]
    assert np.array_equal(bits(sample.cpu().numpy().reshape(-1)), bits(out))
    dec = coder.decode(p, [list(i) for i in indices], seed=seed)
    assert torch.equal(dec, sample)
    # split / merge are inverse of each other and follow the same permutation
    blocks = coder.split(t.loc, seed=seed)[0]
    assert np.array_equal(blocks[3].cpu().numpy(), tl[perm[3000:4000]])
    merged, = coder.merge(blocks, shape=shape, seed=seed)
    assert torch.equal(merged, t.loc)


def test_beam_batch_matches_single(cuda):
    import torch
    from irec_b200 import Normal
    from rec.coding import BeamSearchCoder
    n, N_img = 2048, 5
    coder = BeamSearchCoder(kl_per_partition=3., n_beams=20, extra_samples=1.2, block_size=1000)
    arrs = [synth.c2(n, data_seed=100 + i) for i in range(N_img)]
    stack = [np.stack([a[k] for a in arrs]) for k in range(4)]
    t = Normal(stack[0], stack[1], device=cuda)
    p = Normal(stack[2], stack[3], device=cuda)
    idx_b, samp_b = coder.encode_batch(t, p, seed=42)
    for i in range(N_img):
        ti = Normal(stack[0][i:i + 1], stack[1][i:i + 1], device=cuda)
        pi = Normal(stack[2][i:i + 1], stack[3][i:i + 1], device=cuda)
        idx, samp = coder.encode(ti, pi, seed=42)
        assert idx == idx_b[i]
        assert torch.equal(samp[0], samp_b[i])
    dec = coder.decode_batch(p, idx_b, seed=42)
    assert torch.equal(dec, samp_b)


def test_sharded_block_single_rank(cuda):
    """candidate-range sharded state machine (world size 1) == oracle"""
    from irec_b200 import engine
    mu, sig, pl, ps = synth.c1(64, data_seed=0)
    d = to_dev((mu, sig, pl, ps), cuda)
    S = 2000
    blk = engine.ShardedBeamBlock(64, S, 10, np.log(S) / 1.2, max_aux=64, device=cuda)
    idx, sample = blk.encode(*d, seed=42)
    ref = O.beam_encode_block(mu, sig, pl, ps, np.float32(np.log(S) / 1.2), S, 10, 42)
    assert idx == ref["indices"].tolist()
    assert np.array_equal(bits(sample.cpu().numpy()), bits(ref["sample"]))


@pytest.mark.parametrize("D,S,B,seed", [(64, 2000, 10, 42), (64, 36, 20, 42), (200, 5000, 20, 7), (37, 300, 5, 1), (64, 7, 32, 3), (1, 50, 3, 5)])
def test_fused_block_single_rank(cuda, D, S, B, seed):
    """all auxiliary variables of a block in ONE cooperative launch (irec_beam_encode_fused: state replicas per CTA, last-arriver
    merge of the grid's top-B lists) == oracle == the multi-launch state machine"""
    from irec_b200 import engine
    mu, sig, pl, ps = synth.c1(D, data_seed=D + S)
    d = to_dev((mu, sig, pl, ps), cuda)
    omega = np.float32(np.log(S) / 1.2)
    blk = engine.ShardedBeamBlock(D, S, B, omega, max_aux=64, device=cuda)
    assert blk.fused_available() == (B <= 20)            # 32 beams: the candidate buffer does not fit beside the table
    idx, sample = blk.encode_fused(*d, seed=seed)        # (falls back to the step loop then)
    ref = O.beam_encode_block(mu, sig, pl, ps, omega, S, B, seed)
    assert idx == ref["indices"].tolist()
    assert np.array_equal(bits(sample.cpu().numpy()), bits(ref["sample"]))
    idx2, sample2 = blk.encode(*d, seed=seed)
    assert idx2 == idx and np.array_equal(bits(sample2.cpu().numpy()), bits(sample.cpu().numpy()))


def test_errors(cuda):
    from irec_b200 import Normal
    from rec.coding import BeamSearchCoder
    from rec.coding.utils import CodingError
    coder = BeamSearchCoder(kl_per_partition=3., n_beams=4)
    same = Normal(np.zeros((1, 8)), np.ones((1, 8)), device=cuda)
    with pytest.raises(CodingError):
        coder.encode(same, same, seed=1)                 # KL = 0: the reference crashes, we raise CodingError
    two = Normal(np.zeros((2, 8)), np.ones((2, 8)), device=cuda)
    with pytest.raises(CodingError):
        coder.encode(two, two, seed=1)                   # batch size must be 1
    with pytest.raises(CodingError):
        BeamSearchCoder(3., 4, block_size=4).split(same.loc, Normal(np.zeros((1, 9)), np.ones((1, 9))).loc)
    with pytest.raises(CodingError):
        coder.merge([same.loc.reshape(-1)], shape=None)


# ------------------------------------------------------------------------------------------ importance sampler
@pytest.mark.parametrize("D,coding_bits,seed", [(64, 3 / np.log(2), 42), (1, 8.0, 7), (1000, 5.0, 3), (37, 6.0, 9)])
def test_is_coded_sample(cuda, D, coding_bits, seed):
    import torch
    from irec_b200 import Normal
    from rec.coding.samplers import ImportanceSampler
    tl, ts, pl, ps = synth.c2(D, data_seed=D)
    s = ImportanceSampler(coding_bits=coding_bits)
    t = Normal(tl[None, :], ts[None, :], device=cuda)
    p = Normal(pl[None, :], ps[None, :], device=cuda)
    index, sample = s.coded_sample(t, p, seed)
    oi, osamp = O.is_coded_sample(tl, ts, pl, ps, s.n_samples, seed)
    assert int(index) == oi
    assert np.array_equal(bits(sample.cpu().numpy().reshape(-1)), bits(osamp))
    dec = s.decode_sample(p, index, seed)
    assert torch.equal(dec, sample)


@pytest.mark.parametrize("recipe,D,omega,seed", [("c1", 64, 3.0, 42), ("c2", 1000, 3.0, 5), ("c2", 37, 2.0, 1),
                                                 ("c2", 2000, 3.0, 8)])
def test_gaussian_coder_importance(cuda, recipe, D, omega, seed):
    import torch
    from irec_b200 import Normal
    from rec.coding import GaussianCoder
    from rec.coding.samplers import ImportanceSampler
    tl, ts, pl, ps = getattr(synth, recipe)(D, data_seed=D + 1)
    s = ImportanceSampler(coding_bits=omega / np.log(2), alpha=np.inf)
    coder = GaussianCoder(kl_per_partition=omega, sampler=s)
    t = Normal(tl[None, :], ts[None, :], device=cuda)
    p = Normal(pl[None, :], ps[None, :], device=cuda)
    indices, sample = coder.encode(t, p, seed=seed)
    ref = O.is_encode_block(tl, ts, pl, ps, omega, s.n_samples, seed)
    assert [int(i) for i in indices] == ref["indices"].tolist()
    assert np.array_equal(bits(sample.cpu().numpy().reshape(-1)), bits(ref["sample"]))
    dec = coder.decode(p, list(indices), seed=seed)
    assert torch.equal(dec, sample)
    assert np.array_equal(bits(dec.cpu().numpy().reshape(-1)), bits(O.is_decode_block(pl, ps, seed, ref["indices"])))
    assert abs(coder.get_codelength(indices) - len(indices) * omega) < 1e-4


def test_gaussian_coder_importance_block_size(cuda):
    import torch
    from irec_b200 import Normal
    from rec.coding import GaussianCoder
    from rec.coding.samplers import ImportanceSampler
    n, bs, seed = 2500, 1000, 42
    tl, ts, pl, ps = synth.c2(n, data_seed=33)
    s = ImportanceSampler(coding_bits=3. / np.log(2))
    coder = GaussianCoder(kl_per_partition=3., sampler=s, block_size=bs)
    t = Normal(tl[None, :], ts[None, :], device=cuda)
    p = Normal(pl[None, :], ps[None, :], device=cuda)
    indices, sample = coder.encode(t, p, seed=seed)
    perm = O.shuffle_perm(n, seed)
    out = np.zeros(n, np.float32)
    for b in range(3):
        sel = perm[b * bs:min(n, (b + 1) * bs)]
        ref = O.is_encode_block(tl[sel], ts[sel], pl[sel], ps[sel], 3., s.n_samples, seed)
        assert [int(i) for i in indices[b]] == ref["indices"].tolist()
        out[sel] = ref["sample"]
    assert np.array_equal(bits(sample.cpu().numpy().reshape(-1)), bits(out))
    dec = coder.decode(p, [list(i) for i in indices], seed=seed)
    assert torch.equal(dec, sample)


def test_generic_sampler_plugin_loop(cuda):
    """GaussianCoder with a user-defined Sampler plug-in goes through the reference's Python loop and still
    round-trips (here: a thin subclass that defeats the fused-kernel detection)."""
    import torch
    from irec_b200 import Normal
    from rec.coding import GaussianCoder
    from rec.coding.samplers import ImportanceSampler

    class MySampler(ImportanceSampler):
        pass

    class Wrapper(MySampler):
        @property
        def alpha(self):
            return float("inf")

        @alpha.setter
        def alpha(self, v):
            pass

    mu, sig, pl, ps = synth.c1(64, data_seed=2)
    coder = GaussianCoder(kl_per_partition=3., sampler=Wrapper(coding_bits=3. / np.log(2)))
    coder._uses_kernels = lambda: False
    t = Normal(mu[None, :], sig[None, :], device=cuda)
    p = Normal(pl[None, :], ps[None, :], device=cuda)
    indices, sample = coder.encode(t, p, seed=42)
    dec = coder.decode(p, list(indices), seed=42)
    assert torch.allclose(dec, sample, atol=1e-5)


def test_sharded_block_nccl_two_ranks(cuda):
    """candidate-range sharding over 2 GPUs (NCCL all-gather of top-B records) == oracle; needs >= 2 GPUs"""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29517", os.path.join(root, "tests", "run_sharded_nccl.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "sharded nccl ok" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_p2p_exchange_single_rank_loopback(cuda):
    """k_p2p_exchange with world = 1 (the rank's own buffer is its only peer): records come back unchanged, the device-side
    step counter advances, both parity slots get used.  The multi-rank path is exercised by bench_sweep.py --check under
    torchrun (profiles/r1_sweep_c5_*_p2p_graph.jsonl: matches_oracle)."""
    import ctypes as C
    import torch
    from irec_b200 import native as N
    lib = N.lib()
    B = 20
    words = int(lib.irec_p2p_exchange_bytes(B, 1)) // 4
    buf = torch.zeros(words, dtype=torch.int32, device=cuda)
    ptrs = torch.tensor([buf.data_ptr()], dtype=torch.int64, device=cuda)
    out_rec = torch.zeros(B * 4, dtype=torch.int32, device=cuda)
    out_cnt = torch.zeros(1, dtype=torch.int32, device=cuda)
    for step in range(1, 4):
        local = torch.arange((B + 1) * 4, dtype=torch.int32, device=cuda) * step
        local[B * 4] = 7 + step
        N.check(lib.irec_p2p_exchange(N.ptr(ptrs), 0, 1, B, N.ptr(local), N.ptr(out_rec), N.ptr(out_cnt), N.stream_ptr()),
                "irec_p2p_exchange")
        torch.cuda.synchronize()
        assert torch.equal(out_rec, local[:B * 4]) and int(out_cnt[0]) == 7 + step
        assert int(buf[0]) == step and int(buf[64]) == step


def test_encode_batch_lazy_equals_eager(cuda):
    """encode_batch(lazy=True): index lists read back on demand == the synchronous call, launches may be interleaved"""
    import torch
    from irec_b200 import Normal
    from rec.coding import BeamSearchCoder
    coder = BeamSearchCoder(kl_per_partition=3., n_beams=20, extra_samples=1.2, block_size=500)
    arrs = [synth.c2(1200, data_seed=300 + i) for i in range(3)]
    t = Normal(np.stack([a[0] for a in arrs]), np.stack([a[1] for a in arrs]), device=cuda)
    p = Normal(np.stack([a[2] for a in arrs]), np.stack([a[3] for a in arrs]), device=cuda)
    eager_idx, eager_sample = coder.encode_batch(t, p, seed=42)
    get1, s1 = coder.encode_batch(t, p, seed=42, lazy=True)
    get2, s2 = coder.encode_batch(t, p, seed=43, lazy=True)          # a second launch before the first is read back
    assert get1() == eager_idx and torch.equal(s1, eager_sample)
    other_idx, other_sample = coder.encode_batch(t, p, seed=43)
    assert get2() == other_idx and torch.equal(s2, other_sample)
    assert get1() is get1() or get1() == eager_idx                     # cached


def test_exponent_table_cache_is_transparent(cuda):
    """The per-device cache of the exponent table (rows kept across launches for equal seed / S / block sizes) must never
    change a result: a sequence of launches that hits, extends (more auxiliary variables), misses (other block sizes) and
    invalidates (other seed) the cache gives the same indices and sample bits as with IREC_R2_NO_CACHE=1."""
    import os
    import torch
    from irec_b200 import Normal
    from rec.coding import BeamSearchCoder

    def tensor(n, seed, boost=1.0):
        tl, ts, pl, ps = synth.c2(n, data_seed=seed)
        tl = (pl + boost * (tl - pl)).astype(np.float32)
        return Normal(tl[None], ts[None], device=cuda), Normal(pl[None], ps[None], device=cuda)

    coder = BeamSearchCoder(kl_per_partition=3., n_beams=20, extra_samples=1.2, block_size=1000)
    calls = [(tensor(3192, 1), 42), (tensor(3192, 2), 42), (tensor(3192, 3, boost=1.6), 42),     # hit, then more variables
             (tensor(2500, 4), 42), (tensor(3192, 1), 42),                                          # other sizes, back again
             (tensor(3192, 1), 43), (tensor(3192, 5), 42)]                                          # other seed, back again
    def run():
        out = []
        for (t, p), seed in calls:
            idx, sample = coder.encode(t, p, seed=seed)
            out.append((idx, sample.clone()))
        return out
    os.environ["IREC_R2_NO_CACHE"] = "1"
    try:
        ref = run()
    finally:
        del os.environ["IREC_R2_NO_CACHE"]
    got = run()
    again = run()
    for (ri, rs), (gi, gs), (ai, as_) in zip(ref, got, again):
        assert gi == ri and torch.equal(gs, rs)
        assert ai == ri and torch.equal(as_, rs)


def test_exponent_table_cache_across_streams(cuda):
    """The exponent table is shared by all streams of a device (csrc/irec_beam.cu, r2_tab_acquire): launches of different
    streams read it concurrently and append rows concurrently; a launch with OTHER block sizes may rewrite it only when no
    other stream can be reading, and otherwise builds a private table in its workspace.  Launches of two streams are
    enqueued back to back without any host synchronisation -- same sizes (shared reads), more auxiliary variables on one
    stream (append under a reader), other sizes (private table / re-key) -- and every result must equal IREC_R2_NO_CACHE=1."""
    import os
    import torch
    from irec_b200 import Normal
    from rec.coding import BeamSearchCoder

    def batch(n_img, n, seed, boost=1.0):
        arrs = [synth.c2(n, data_seed=seed + i) for i in range(n_img)]
        tl, ts, pl, ps = (np.stack([a[k] for a in arrs]) for k in range(4))
        tl = (pl + boost * (tl - pl)).astype(np.float32)
        return Normal(tl, ts, device=cuda), Normal(pl, ps, device=cuda)

    coder = BeamSearchCoder(kl_per_partition=3., n_beams=20, extra_samples=1.2, block_size=1000)
    a1, a2, a3 = batch(80, 3192, 100), batch(80, 3192, 300), batch(80, 3192, 500, boost=1.7)      # 320 coder-blocks each
    b1, b2 = batch(60, 2500, 700), batch(60, 2700, 900)                                            # other block sizes
    # (stream, inputs): consecutive entries overlap on the device
    plan = [(0, a1), (1, a2), (0, a1), (1, a3), (0, a2), (1, b1), (0, a1), (1, b2), (0, b1), (1, a3), (0, b2), (1, b1)]

    def run(streams):
        pend = []
        for k, (t, p) in plan:
            with torch.cuda.stream(streams[k]):
                pend.append(coder.encode_batch(t, p, seed=42, lazy=True))
        return [(get(), smp) for get, smp in pend]

    os.environ["IREC_R2_NO_CACHE"] = "1"
    try:
        cur = torch.cuda.current_stream()
        ref = run([cur, cur])
        torch.cuda.synchronize()
    finally:
        del os.environ["IREC_R2_NO_CACHE"]
    s = [torch.cuda.Stream(), torch.cuda.Stream()]
    for st in s:
        st.wait_stream(torch.cuda.current_stream())
    for rep in range(2):
        got = run(s)
        torch.cuda.synchronize()
        for i, ((ri, rs), (gi, gs)) in enumerate(zip(ref, got)):
            assert gi == ri, f"launch {i} of repetition {rep}: indices differ"
            assert torch.equal(gs, rs), f"launch {i} of repetition {rep}: sample differs"


def test_reserved_sms_do_not_change_results(cuda):
    """irec_set_thread_reserved_sms: the persistent batch kernel runs on fewer SMs (any k, clamped to leave one CTA); same
    indices and sample bits"""
    import torch
    from irec_b200 import Normal
    from irec_b200.native import reserved_sms
    from rec.coding import BeamSearchCoder
    coder = BeamSearchCoder(kl_per_partition=3., n_beams=20, extra_samples=1.2, block_size=1000)
    arrs = [synth.c2(8192, data_seed=800 + i) for i in range(40)]                   # 360 coder-blocks: the tensor-memory kernel
    t = Normal(np.stack([a[0] for a in arrs]), np.stack([a[1] for a in arrs]), device=cuda)
    p = Normal(np.stack([a[2] for a in arrs]), np.stack([a[3] for a in arrs]), device=cuda)
    ref_idx, ref_sample = coder.encode_batch(t, p, seed=42)
    for k in (4, 100, 10 ** 6):
        with reserved_sms(k):
            idx, sample = coder.encode_batch(t, p, seed=42)
        assert idx == ref_idx and torch.equal(sample, ref_sample), k
    idx, sample = coder.encode_batch(t, p, seed=42)                                # setting cleared on exit
    assert idx == ref_idx and torch.equal(sample, ref_sample)


# ------------------------------------------------------------------------------------------ round 2: decode validation, IS candidate table
def test_decode_rejects_index_lists_beyond_the_ratio_table(cuda):
    """a learned ratio table shorter than an index list must raise CodingError (reference coder.py:226-231),
    not read past the table; the kernels flag such rows and write NaN instead of garbage."""
    import torch
    from irec_b200 import Normal, engine as E, native as N
    from rec.coding import BeamSearchCoder, GaussianCoder
    from rec.coding.samplers import ImportanceSampler
    from rec.coding.utils import CodingError
    tl, ts, pl, ps = synth.c2(100, data_seed=3)
    p = Normal(pl[None, :], ps[None, :], device=cuda)
    table = torch.tensor([1.0, 0.5, 0.4], dtype=torch.float32, device=cuda)
    long_list = [1, 2, 3, 4, 5]
    for coder in (BeamSearchCoder(kl_per_partition=3., n_beams=4), GaussianCoder(3., ImportanceSampler(coding_bits=4.))):
        with N.thread_aux_ratios(table):
            assert N.load_library().irec_aux_ratio_len() == 3
            with pytest.raises(CodingError):
                coder._decode_flat(p.loc.reshape(-1), p.scale.reshape(-1), None, *E.make_block_offsets(100, None, cuda)[:2], 100, 1,
                                   [list(long_list)])
        assert N.load_library().irec_aux_ratio_len() == 65536
    # device-side check (index counts only known on the device): status + NaN
    offs, nb, _ = E.make_block_offsets(200, 100, cuda)
    pl2 = torch.zeros(200, device=cuda); ps2 = torch.ones(200, device=cuda)
    idx = torch.zeros((2, 8), dtype=torch.int32, device=cuda)
    n_aux = torch.tensor([2, 7], dtype=torch.int32, device=cuda)
    with N.thread_aux_ratios(table):
        out, status = E.beam_decode_blocks(pl2, ps2, None, offs, nb, 20, 1, (idx, n_aux, 8), return_status=True)
    assert status.tolist() == [N.BLK_OK, N.BLK_TOO_LONG]
    assert torch.isfinite(out[:100]).all() and torch.isnan(out[100:]).all()
    with pytest.raises(IndexError):                       # reference: indices[0] of an empty list
        E.is_decode_blocks(pl2, ps2, None, offs, nb, 100, 1, [[1], []])


@pytest.mark.parametrize("n,bs", [(2500, 1000), (300, 128), (37, None)])
def test_is_encode_candidate_table_equals_in_place(cuda, n, bs):
    """the launch-wide candidate table (k_is_ztab) and in-place Philox/Box-Muller give the same indices and sample bits"""
    import torch
    from irec_b200 import Normal
    from rec.coding import GaussianCoder
    from rec.coding.samplers import ImportanceSampler
    tl, ts, pl, ps = synth.c2(n, data_seed=91)
    t = Normal(tl[None, :], ts[None, :], device=cuda)
    p = Normal(pl[None, :], ps[None, :], device=cuda)
    res = []
    for no_table in ("0", "1"):
        os.environ["IREC_IS_NO_TABLE"] = no_table
        try:
            coder = GaussianCoder(kl_per_partition=3., sampler=ImportanceSampler(coding_bits=3. / np.log(2)), block_size=bs)
            res.append(coder.encode(t, p, seed=11))
        finally:
            os.environ.pop("IREC_IS_NO_TABLE", None)
    assert res[0][0] == res[1][0]
    assert torch.equal(res[0][1], res[1][1])


def test_encode_grows_row_capacity_without_presizing(cuda):
    """no sizing pre-pass: a block that needs more auxiliary variables than the guessed capacity is re-launched with the
    exact one (same result as with a promised max_aux)"""
    import torch
    from irec_b200 import Normal, engine as E
    from rec.coding import BeamSearchCoder
    tl, ts, pl, ps = synth.c2(1000, data_seed=5)
    coder = BeamSearchCoder(kl_per_partition=1.0, n_beams=3, extra_samples=2.0)     # ~260 nats / 1 nat: > 128 variables
    t = Normal(tl[None, :], ts[None, :], device=cuda)
    p = Normal(pl[None, :], ps[None, :], device=cuda)
    E._aux_hint.clear()
    idx1, s1 = coder.encode(t, p, seed=3)
    assert len(idx1) > E._AUX_HINT_FIRST
    ref = O.beam_encode_block(tl, ts, pl, ps, 1.0, coder.n_samples, 3, 3)
    assert list(idx1) == ref["indices"].tolist()
    assert np.array_equal(bits(s1.cpu().numpy().reshape(-1)), bits(ref["sample"]))
    idx2, s2 = coder.encode(t, p, seed=3)                 # the hint now covers it: single launch
    assert idx2 == idx1 and torch.equal(s1, s2)
    E._aux_hint.clear()
    get, s3 = coder.encode_batch(t, p, seed=3, lazy=True)
    assert get()[0] == idx1 and torch.equal(s3.reshape(-1), s1.reshape(-1))


def test_schedule_export_and_pseudo_random_sample(cuda):
    """irec_schedule == the oracle's schedule bit for bit, and BeamSearchCoder.get_pseudo_random_sample
    (reference beam_search_coder.py:37-51) == the reference-structured port's candidate tensor"""
    import torch
    from irec_b200 import Normal, engine as E
    from oracle import ref_numpy as R
    from rec.coding import BeamSearchCoder
    for recipe, D in (("c2", 1000), ("c3", 288), ("c1", 64), ("c2", 37)):
        tl, ts, pl, ps = getattr(synth, recipe)(D, data_seed=4)
        d = to_dev((tl, ts, pl, ps), cuda)
        kl, n, sa, A, Ec, M = E.schedule(*d, 3.0)
        okl = O.kl(tl, ts, pl, ps)
        assert bits(np.float32(kl)) == bits(np.float32(okl)) and n == O.n_aux(okl, 3.0)
        osa, oA, oE, oM = O.beam_schedule(tl, ts, pl, ps, n)
        for got, ref in ((sa, osa), (A, oA), (Ec, oE), (M, oM)):
            assert np.array_equal(bits(got.cpu().numpy()), bits(np.asarray(ref, np.float32).reshape(n, D)))
    coder = BeamSearchCoder(kl_per_partition=3., n_beams=20, extra_samples=1.2)
    scale = np.exp(np.random.default_rng(1).uniform(-1, 0, 50)).astype(np.float32)
    idx = np.array([[3, 5], [7, 1], [0, 35]], np.int32)
    got = coder.get_pseudo_random_sample(Normal(np.zeros((1, 50), np.float32), scale[None, :], device=cuda), coder.n_samples, idx, 43)
    ref = R._pseudo_random_sample(scale, coder.n_samples, idx, 43)
    assert got.shape == (36, 3, 50)
    assert np.array_equal(bits(got.cpu().numpy()), bits(ref))

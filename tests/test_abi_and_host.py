"""CPU tests: libirec.so builds, loads and exports every symbol include/irec.h declares (no compute calls),
plus the host-side logic that lives in the library (TF op seeds, Coder.split permutation) against the oracle."""
import os
import random
import re

import numpy as np
import pytest

from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    from irec_b200 import native
    lib = native.load_library()
    header = open(os.path.join(ROOT, "include", "irec.h")).read() + open(os.path.join(ROOT, "include", "irec_io.h")).read()
    declared = set(re.findall(r"\b(irec_[a-z0-9_]+)\s*\(", header))
    declared -= {"irec_record_t", "irec_rec_header_t"}
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"libirec.so does not export {name}"
    assert declared == set(native.EXPORTED_SYMBOLS), declared ^ set(native.EXPORTED_SYMBOLS)
    assert lib.irec_version() >= 100


def test_no_cpu_fallback(built):
    """without a GPU the product path must fail loudly, not fall back"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from irec_b200 import Normal, native
    from rec.coding import BeamSearchCoder
    coder = BeamSearchCoder(kl_per_partition=3., n_beams=2)
    with pytest.raises(native.NativeError):
        coder.encode(Normal(np.zeros((1, 4)), np.ones((1, 4))), Normal(np.zeros((1, 4)), np.ones((1, 4))), seed=1)



def test_tf_op_seed_host(built):
    from irec_b200 import native
    for s in (0, 1, 42, 43, 69420, 2 ** 31, 2 ** 40 + 5):
        assert native.tf_op_seed(s) == random.Random(s).randint(0, 2 ** 31 - 1) == O.py_randint31(s)


def test_aux_ratio_host(built):
    from irec_b200 import native
    for i in (0, 1, 2, 10, 99, 1000):
        assert np.float32(native.aux_ratio(i)) == np.float32(np.power(i + 1., -0.7864636765648174))
        assert np.float32(native.aux_ratio(i)) == np.float32(O.aux_ratio(i))


@pytest.mark.parametrize("n,seed", [(1, 42), (2, 42), (10, 42), (8192, 42), (12288, 7), (301056, 42)])
def test_split_permutation_matches_oracle(built, n, seed):
    from irec_b200 import native
    p = native.split_permutation(n, seed).numpy()
    assert np.array_equal(p, O.shuffle_perm(n, seed))
    assert np.array_equal(np.sort(p), np.arange(n))


def test_api_surface(built):
    """the names and keyword signatures the reference's callers use"""
    import inspect
    from rec.coding import Coder, GaussianCoder, BeamSearchCoder
    from rec.coding.samplers import ImportanceSampler, RejectionSampler, Sampler
    from rec.coding.utils import CodingError
    from rec.coding.importance_sampling import encode_gaussian_importance_sample, decode_gaussian_importance_sample
    b = BeamSearchCoder(kl_per_partition=3., n_beams=20, extra_samples=1.2, name="x", block_size=1000)
    assert b.n_samples == 36 and b.n_beams == 20 and b.block_size == 1000 and b.big_prime == 10007
    assert abs(b.get_codelength([1, 2, 3]) - 3 * np.log(36)) < 1e-12
    s = ImportanceSampler(coding_bits=3 / np.log(2), alpha=np.inf, extrapolate_auxiliary_vars=True)
    g = GaussianCoder(kl_per_partition=3., sampler=s, block_size=1000)
    assert abs(g.get_codelength([0, 0]) - 6.0) < 1e-5
    assert list(inspect.signature(BeamSearchCoder.encode).parameters)[:4] == ["self", "target_dist", "coding_dist", "seed"]
    assert list(inspect.signature(BeamSearchCoder.decode).parameters)[:4] == ["self", "coding_dist", "indices", "seed"]
    assert issubclass(BeamSearchCoder, GaussianCoder) and issubclass(GaussianCoder, Coder)
    assert issubclass(CodingError, Exception)
    assert b.simple_hash([[3, 5]]).tolist() == [O.simple_hash([3, 5])]
    r = RejectionSampler(sample_buffer_size=10, r_buffer_size=20)      # reference samplers.py:106
    assert isinstance(r, Sampler) and r.sample_buffer_size == 10 and r.r_buffer_size == 20
    from rec.coding.sample_generator import PseudoSampleGenerator
    assert isinstance(RejectionSampler(sample_buffer_size=10, r_buffer_size=20, use_pseudo_sampler=True).sample_generator,
                      PseudoSampleGenerator)
    with pytest.raises(CodingError):
        encode_gaussian_importance_sample(None, None, None, None, 3., 1, alpha=0.5)
    with pytest.raises(CodingError):
        GaussianCoder(kl_per_partition=3., sampler=s, extrapolate_auxiliary_ratios=False).get_auxiliary_ratio(0)


def test_importance_num_samples_matches_oracle(built):
    from rec.coding.importance_sampling import importance_num_samples
    for cb in (1.0, 3 / np.log(2), 5.5, 10.0, 16.0, 20.0):
        assert importance_num_samples(cb) == O.is_num_samples(cb)


def test_product_does_not_import_oracle():
    """the oracle is test infrastructure: nothing under the product package may import or link it"""
    pkg = os.path.join(ROOT, "relative-entropy-coding_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                for line in open(os.path.join(dp, f)).read().splitlines():
                    low = line.lower()
                    assert not (("import" in low or "include" in low or "cdll" in low) and "oracle" in low), (f, line)


def test_distributions_accept_dlpack_producers(built):
    """SURVEY.md 8b (data types): distributions are duck-typed (.loc / .scale); values may be torch tensors, array-likes or
    any `__dlpack__` producer (a tf eager tensor, a cupy / jax array) -- taken over without a copy through DLPack"""
    import numpy as np
    import torch
    from irec_b200 import Normal

    class Producer:                                   # only the DLPack protocol, nothing else
        def __init__(self, a):
            self._a = a

        def __dlpack__(self, **kw):
            return self._a.__dlpack__(**kw)

        def __dlpack_device__(self):
            return self._a.__dlpack_device__()

    loc = np.arange(12, dtype=np.float32).reshape(1, 12)
    scale = np.full((1, 12), 0.5, np.float32)
    d = Normal(Producer(loc), Producer(scale))
    assert isinstance(d.loc, torch.Tensor) and d.loc.dtype == torch.float32
    assert np.array_equal(d.loc.numpy(), loc) and np.array_equal(d.scale.numpy(), scale)
    loc[0, 0] = 7.0                                   # zero-copy: the holder sees the producer's memory
    assert float(d.loc[0, 0]) == 7.0


def test_reserved_sms_setting_is_host_only(built):
    """irec_set_thread_reserved_sms is a per-thread host setting (no device call): usable before any CUDA context exists"""
    from irec_b200 import native as N
    lib = N.load_library()
    assert lib.irec_set_thread_reserved_sms(4) == 0
    assert lib.irec_set_thread_reserved_sms(-1) == 0
    with N.reserved_sms(8):
        pass

"""Index-stream wire format: arithmetic coder + `.rec` container.

Three parties on the same inputs: the product (host C++ in libirec.so behind rec.io), the Python restatement
(oracle/io_port.py) and the reference ITSELF -- golden vectors its compiled rec/io wrote (tests/golden/io_golden.json)
and, where oracle/_ref is present, the live reference in a subprocess.  Everything here is integer/byte work: bit-exact.
CPU tests (the reference's IO is CPU code; the product's is host C++ by design)."""
import json
import os

import numpy as np
import pytest

from oracle import io_port, ref_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "io_golden.json")))


@pytest.fixture(scope="module")
def io(built):
    import rec.io as rio
    return rio


@pytest.mark.parametrize("k", range(len(GOLD["ac"])))
def test_port_matches_reference_golden_codes(k):
    c = GOLD["ac"][k]
    assert io_port.ac_encode(c["P"], c["message"]) == c["code"]
    assert io_port.ac_decode(c["P"], c["code"]) == c["message"]


@pytest.mark.parametrize("k", range(len(GOLD["ac"])))
def test_arithmetic_coder_matches_reference_golden_codes(io, k):
    c = GOLD["ac"][k]
    ac = io.ArithmeticCoder(np.array(c["P"], dtype=np.int32), precision=32)
    code = ac.encode(np.array(c["message"]))
    assert isinstance(code, list) and "".join(code) == c["code"]
    assert ac.decode_fast(code) == c["message"]
    assert ac.decode(c["code"]) == c["message"]
    assert ac.R == sum(c["P"]) and list(ac.C[:2]) == [0, c["P"][0]]


@pytest.mark.parametrize("k", range(len(GOLD["rec"])))
def test_container_matches_reference_golden_files(io, k, tmp_path):
    c = GOLD["rec"][k]
    gold = bytes.fromhex(c["file_hex"])
    assert io_port.rec_pack(c["seed"], c["image_shape"], c["block_size"], c["block_indices"], c["max_index"]) == gold
    path = str(tmp_path / "x.rec")
    n = io.write_compressed_code(path, c["seed"], tuple(c["image_shape"]), c["block_size"], c["block_indices"], c["max_index"])
    data = open(path, "rb").read()
    assert n == len(data) and data == gold                       # byte-identical to the reference's file
    gpath = str(tmp_path / "gold.rec")
    open(gpath, "wb").write(gold)                                # and the reference's file reads back
    seed, shape, bs, bi = io.read_compressed_code(gpath)
    assert (seed, list(shape), bs, bi) == (c["seed"], c["image_shape"], c["block_size"], c["block_indices"])
    assert io_port.rec_unpack(gold)[3] == c["block_indices"]


def test_coder_random_roundtrips_against_port(io):
    """the shape of the reference's own script test (rec/io/tests/coding_test.py:9-46), many seeds, vs the restatement"""
    rng = np.random.Generator(np.random.PCG64(7))
    for trial in range(40):
        n_sym = int(rng.integers(1, 70))
        P = np.ones(n_sym + 1, dtype=np.int32)
        P[1:] = rng.integers(1, 101, n_sym)
        n = int(rng.integers(0, 400))
        msg = np.zeros(n + 1, dtype=np.int32)
        msg[:-1] = rng.integers(1, n_sym + 1, n)
        ac = io.ArithmeticCoder(P, precision=32)
        code = "".join(ac.encode(msg))
        assert code == io_port.ac_encode(P.tolist(), msg.tolist())
        assert ac.decode_fast(code) == msg.tolist()


def test_coder_full_size_roundtrip(io):
    """coding_test.py sizes: 64 symbols, 2000-symbol message; code length within 2 bits + 1% of the entropy bound"""
    rng = np.random.Generator(np.random.PCG64(11))
    P = np.ones(65, dtype=np.int32)
    P[1:] = rng.integers(1, 101, 64)
    msg = np.zeros(2000, dtype=np.int32)
    msg[:-1] = rng.integers(1, 65, 1999)
    ac = io.ArithmeticCoder(P, precision=32)
    code = ac.encode(msg)
    assert ac.decode_fast(code) == msg.tolist()
    ideal = -np.log2(P[msg] / P.sum()).sum()
    assert ideal <= len(code) <= ideal * 1.01 + 4


def test_container_errors(io, tmp_path):
    from irec_b200.native import NativeError
    with pytest.raises(ValueError):
        io.write_compressed_code(str(tmp_path / "a.rec"), 1, (32, 32), 1000, [[[1]]], 20)
    with pytest.raises(NativeError):          # index 35 with max_index 20: the reference dies with an IndexError here
        io.write_compressed_code(str(tmp_path / "a.rec"), 1, (32, 32, 3), 1000, [[[35, 2]]], 20)
    with pytest.raises(NativeError):
        open(tmp_path / "short.rec", "wb").write(b"\x00" * 10)
        io.read_compressed_code(str(tmp_path / "short.rec"))
    ac = io.ArithmeticCoder(np.array([1, 5, 5]))
    with pytest.raises(NativeError):
        ac.encode([3, 0])                      # symbol outside the alphabet
    assert ac.decode_fast("") == [0]           # z = 0 lies in the end symbol's interval (the reference answers the same)
    assert ac.decode_fast("1" * 40) == io_port.ac_decode([1, 5, 5], "1" * 40)   # bits past the end read as zeros


@pytest.mark.skipif(not ref_io.available(), reason="oracle/_ref (the compiled reference rec/io) is not built")
def test_live_reference_agrees(io, tmp_path):
    """fresh random inputs through the real reference, the port and the product"""
    rng = np.random.Generator(np.random.PCG64(123))
    reqs, cases = [], []
    for trial in range(12):
        n_sym = int(rng.integers(1, 50))
        P = [1] + [int(v) for v in rng.integers(1, 2000, n_sym)]
        msg = [int(v) for v in rng.integers(1, n_sym + 1, int(rng.integers(0, 300)))] + [0]
        cases.append((P, msg))
        reqs.append({"op": "ac_encode", "P": P, "message": msg})
    bi = [[[int(v) for v in rng.integers(0, 20, int(rng.integers(0, 30)))] for _ in range(int(rng.integers(1, 12)))]
          for _ in range(5)]
    ref_path, my_path = str(tmp_path / "ref.rec"), str(tmp_path / "mine.rec")
    reqs.append({"op": "rec_write", "path": ref_path, "seed": 7, "image_shape": [64, 48, 3], "block_size": 1000,
                 "block_indices": bi, "max_index": 20})
    io.write_compressed_code(my_path, 7, (64, 48, 3), 1000, bi, 20)
    reqs.append({"op": "rec_read", "path": my_path})             # the reference reads OUR file
    replies = ref_io.call(reqs)
    for (P, msg), code in zip(cases, replies[:len(cases)]):
        assert "".join(io.ArithmeticCoder(np.array(P)).encode(msg)) == code == io_port.ac_encode(P, msg)
        assert io.ArithmeticCoder(np.array(P)).decode_fast(code) == msg
    assert open(ref_path, "rb").read() == open(my_path, "rb").read()
    assert replies[-1]["block_indices"] == bi and replies[-1]["image_shape"] == [64, 48, 3]
    assert io.read_compressed_code(ref_path) == (7, (64, 48, 3), 1000, bi)

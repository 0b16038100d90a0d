"""CPU restatement of the reference's index-stream wire format (test infrastructure only; nothing under the product
package imports it).  Pinned against the reference itself: tests/golden/io_golden.json was written by the real
rec/io code compiled into oracle/_ref (tests/golden/make_io_golden.py), and tests/test_io_parity.py compares this port,
the C++ product code and -- where oracle/_ref is present -- the live reference on the same inputs.

  ac_encode / ac_decode : rec/io/entropy_coding.pyx:51-117, 212-302 (Python integers, so exact)
  rec_pack / rec_unpack : rec/io/utils.py:7-106, 109-216
"""
import struct


def _model(P):
    C, D, c = [], [], 0
    for p in P:                     # entropy_coding.pyx:33-41
        C.append(c)
        c += int(p)
        D.append(c)
    return C, D, c


def ac_encode(P, message, precision=32):
    C, D, R = _model(P)
    whole, half, quarter = 2 ** precision, 2 ** (precision - 1), 2 ** (precision - 2)
    low, high, s, code = 0, whole, 0, []
    for m in message:
        width = high - low
        high = low + (width * D[m]) // R
        low = low + (width * C[m]) // R
        while high < half or low > half:            # :84-101
            if high < half:
                code.append("0" + "1" * s)
                s = 0
                low *= 2
                high *= 2
            elif low > half:
                code.append("1" + "0" * s)
                s = 0
                low = (low - half) * 2
                high = (high - half) * 2
        while low > quarter and high < 3 * quarter:  # :104-107
            s += 1
            low = (low - quarter) * 2
            high = (high - quarter) * 2
    s += 1
    code.append(("0" + "1" * s) if low <= quarter else ("1" + "0" * s))
    return "".join(code)


def ac_decode(P, code, precision=32):
    C, D, R = _model(P)
    whole, half, quarter = 2 ** precision, 2 ** (precision - 1), 2 ** (precision - 2)
    low, high, z, i = 0, whole, 0, 0
    while i < precision and i < len(code):
        if code[i] == "1":
            z += 2 ** (precision - i - 1)
        i += 1
    message = []
    while True:
        width = high - low
        j = max(k for k in range(len(C)) if (width * C[k]) // R <= z - low)    # tightest lower bound (data_structures.py:184-210)
        low, high = low + (width * C[j]) // R, low + (width * D[j]) // R
        message.append(j)
        if j == 0:
            return message
        while high < half or low > half:
            if high < half:
                low, high, z = low * 2, high * 2, z * 2
            elif low > half:
                low, high, z = (low - half) * 2, (high - half) * 2, (z - half) * 2
            if i < len(code) and code[i] == "1":
                z += 1
            i += 1
        while low > quarter and high < 3 * quarter:
            low, high, z = (low - quarter) * 2, (high - quarter) * 2, (z - quarter) * 2
            if i < len(code) and code[i] == "1":
                z += 1
            i += 1


def _to_bytes(code):
    c = "1" + code                  # leading 1 keeps leading zeros (utils.py:64-72)
    n = len(c) // 8 + (1 if len(c) % 8 else 0)
    return int(c, 2).to_bytes(length=n, byteorder="big")


def rec_pack(seed, image_shape, block_size, block_indices, max_index):
    h, w, c = image_shape
    n = len(block_indices)
    index_counts = [1] + [1001] * max_index
    navs = [[len(b) for b in blk] for blk in block_indices]
    nav_codes, idx_codes, maxes = [], [], []
    for nav, blk in zip(navs, block_indices):
        mx = max(nav)
        maxes.append(mx)
        nav_codes.append(_to_bytes(ac_encode([1] + [101] * (mx + 1), [v + 1 for v in nav] + [0])))
        flat = [int(v) for b in blk for v in b]
        idx_codes.append(_to_bytes(ac_encode(index_counts, [v + 1 for v in flat] + [0])))
    header = struct.pack(f"IIIIIHHHH{n}I{n}I{n}I{n}I", seed, block_size, max_index, h, w, c, 0, 0, n,
                         *[len(blk) for blk in block_indices], *[len(b) for b in nav_codes], *[len(b) for b in idx_codes], *maxes)
    return header + b"".join(nav_codes) + b"".join(idx_codes)


def rec_unpack(data):
    seed, block_size, max_index, h, w, c, f_nav, f_idx, n = struct.unpack("IIIIIHHHH", data[:28])
    dyn = struct.unpack(f"{4 * n}I", data[28:28 + 16 * n])
    num_blocks, nav_len, idx_len, maxes = dyn[:n], dyn[n:2 * n], dyn[2 * n:3 * n], dyn[3 * n:]
    pos = 28 + 16 * n
    nav_codes, idx_codes = [], []
    for ln in nav_len:
        nav_codes.append(bin(int.from_bytes(data[pos:pos + ln], "big"))[3:])
        pos += ln
    for ln in idx_len:
        idx_codes.append(bin(int.from_bytes(data[pos:pos + ln], "big"))[3:])
        pos += ln
    index_counts = [1] + [1001] * max_index
    out = []
    for mx, nc, ic in zip(maxes, nav_codes, idx_codes):
        nav = [v - 1 for v in ac_decode([1] + [101] * (mx + 1), nc)[:-1]]
        flat = [v - 1 for v in ac_decode(index_counts, ic)[:-1]]
        blocks, p = [], 0
        for k in nav:
            blocks.append(flat[p:p + k])
            p += k
        out.append(blocks)
    return seed, (h, w, c), block_size, out

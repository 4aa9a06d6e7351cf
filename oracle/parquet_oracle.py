"""CPU ORACLE (test infrastructure): Parquet page payload decoding restated from the reference's vectorized reader.

Follows sql/core/src/main/java/org/apache/spark/sql/execution/datasources/parquet/VectorizedRleValuesReader.java
(:95-117 initFromPage -- bit width byte for values, bitWidth 0 = all zeros; :940-975 readUnsignedVarInt /
readIntLittleEndianPaddedOnBitWidth; :981-1020 readNextGroup -- header & 1 ? bit-packed groups of 8 : RLE run) and
VectorizedColumnReader.java:412-460 (readPageV1: length-prefixed RLE definition levels in front of the values;
readPageV2: levels in their own section).  Pinned in tests/test_scan_cpu.py against pages written by pyarrow (parquet-cpp),
an independent implementation of the format.
"""
import numpy as np


def read_uvarint(b, pos):
    v, shift = 0, 0
    while True:
        x = int(b[pos]); pos += 1
        v |= (x & 0x7F) << shift
        if not (x & 0x80):
            return v, pos
        shift += 7


def decode_hybrid(buf, bit_width, count):
    """RLE / bit-packed hybrid -> `count` unsigned ints (numpy uint32)."""
    out = np.zeros(count, np.uint32)
    if bit_width == 0:
        return out
    b = np.asarray(buf, dtype=np.uint8)
    vbytes = (bit_width + 7) // 8
    pos, n = 0, 0
    while n < count and pos < len(b):
        header, pos = read_uvarint(b, pos)
        if header & 1:
            groups = header >> 1
            nbytes = groups * bit_width
            bits = np.unpackbits(b[pos:pos + nbytes], bitorder="little")
            vals = bits[: groups * 8 * bit_width].reshape(-1, bit_width)
            weights = (1 << np.arange(bit_width, dtype=np.uint64))
            v = (vals.astype(np.uint64) * weights).sum(axis=1).astype(np.uint32)
            take = min(len(v), count - n)
            out[n:n + take] = v[:take]
            n += take
            pos += nbytes
        else:
            run = header >> 1
            val = int.from_bytes(bytes(b[pos:pos + vbytes]), "little")
            pos += vbytes
            take = min(run, count - n)
            out[n:n + take] = val
            n += take
    return out


_NP_PHYS = {0: None, 1: np.dtype("<i4"), 2: np.dtype("<i8"), 4: np.dtype("<f4"), 5: np.dtype("<f8")}


def decode_column_chunk(data, pages, dict_offset, dict_count, physical):
    """pages: [(encoding, num_values, values_offset, values_bytes, def_offset, def_bytes)] -> (values ndarray in ROW space
    with 0 at NULL rows, valid bool ndarray or None)."""
    data = np.asarray(data, dtype=np.uint8)
    dt = _NP_PHYS[physical]
    dictionary = None
    if dict_offset >= 0:
        if dt is None:
            dictionary = data[dict_offset:dict_offset + dict_count].astype(np.uint8)
        else:
            dictionary = np.frombuffer(data[dict_offset:dict_offset + dict_count * dt.itemsize].tobytes(), dtype=dt)
    vals, valids, nullable = [], [], False
    for enc, n, voff, vbytes, doff, dbytes in pages:
        if dbytes > 0:
            valid = decode_hybrid(data[doff:doff + dbytes], 1, n).astype(bool)
            nullable = True
        else:
            valid = np.ones(n, bool)
        nvals = int(valid.sum())
        body = data[voff:voff + vbytes]
        if enc == 0:
            if dt is None:
                dense = np.unpackbits(body, bitorder="little")[:nvals].astype(np.uint8)
            else:
                dense = np.frombuffer(body[: nvals * dt.itemsize].tobytes(), dtype=dt)
        elif enc == 2:      # RLE-encoded BOOLEAN values: [4-byte length][hybrid, bit width 1]
            dense = decode_hybrid(body[4:], 1, nvals).astype(np.uint8)
        elif nvals == 0:    # a page of NULLs only (its dictionary may be empty)
            dense = np.zeros(0, dt or np.uint8)
        else:
            idx = decode_hybrid(body[1:], int(body[0]), nvals)
            dense = dictionary[idx]
        row = np.zeros(n, dense.dtype)
        row[valid] = dense
        vals.append(row)
        valids.append(valid)
    values = np.concatenate(vals) if vals else np.zeros(0, dt or np.uint8)
    return values, (np.concatenate(valids) if nullable else None)

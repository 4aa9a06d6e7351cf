"""FileSourceScanExec mirror for Parquet: column chunks go to the GPU ENCODED and are decoded there (csrc/scan.cu).

Reference: FileSourceScanExec.doExecuteColumnar (sql/core/src/main/scala/org/apache/spark/sql/execution/DataSourceScanExec.scala:735-760)
hands out the ColumnarBatches of VectorizedParquetRecordReader (sql/core/src/main/java/.../parquet/VectorizedParquetRecordReader.java);
one batch here is one row group.  The footer (schema, column-chunk byte ranges) is read with pyarrow's metadata reader -- the
role parquet-mr's ParquetFileReader plays for the reference; every data byte goes through sb_parquet_chunk_pages +
sb_scan_decode.
"""
from __future__ import annotations

import ctypes as C
import mmap

import numpy as np
import pyarrow as pa
import pyarrow.parquet as pq

from . import _capi as capi
from .columnar import ColumnarBatch, _h
from .execution import SparkPlan

_PHYS = {"BOOLEAN": capi.SB_PHYS["BOOLEAN"], "INT32": capi.SB_PHYS["INT32"], "INT64": capi.SB_PHYS["INT64"],
         "FLOAT": capi.SB_PHYS["FLOAT"], "DOUBLE": capi.SB_PHYS["DOUBLE"]}


def _sb_type(arrow_type):
    if pa.types.is_decimal(arrow_type):
        if arrow_type.precision > 18:
            raise capi.SparkB200Error(5, "decimal precision %d > 18 is not supported by the GPU scan" % arrow_type.precision)
        return capi.SB_DECIMAL64, arrow_type.scale
    from .columnar import sb_type_of
    return sb_type_of(arrow_type), 0


class EncodedColumnChunk:
    """Host image of one column chunk + its page descriptors (keeps the buffers alive for the C struct)."""

    def __init__(self, sb_type, scale, physical, data: np.ndarray, pages, npages, dict_offset, dict_count):
        self.type, self.scale, self.physical = sb_type, scale, physical
        self.data, self.pages, self.npages = data, pages, npages
        self.dict_offset, self.dict_count = dict_offset, dict_count

    @property
    def nbytes(self):
        return int(self.data.nbytes)

    def c(self) -> capi.sb_column_chunk:
        s = capi.sb_column_chunk()
        s.type, s.scale, s.physical_type, s.npages = self.type, self.scale, self.physical, self.npages
        s.data = self.data.ctypes.data
        s.data_bytes = self.data.nbytes
        s.pages = C.cast(self.pages, C.POINTER(capi.sb_page))
        s.dict_offset, s.dict_count = self.dict_offset, self.dict_count
        return s

    @staticmethod
    def from_file_bytes(buf: np.ndarray, sb_type, scale, physical, max_def_level, pages_cap=None) -> "EncodedColumnChunk":
        lib = capi.load()
        cap = pages_cap or max(16, buf.nbytes // 512 + 16)
        pages = (capi.sb_page * cap)()
        npages, doff, dcount = C.c_int32(), C.c_int64(), C.c_int32()
        capi.check(lib.sb_parquet_chunk_pages(buf.ctypes.data, buf.nbytes, max_def_level, pages, cap, C.byref(npages), C.byref(doff), C.byref(dcount)))
        return EncodedColumnChunk(sb_type, scale, physical, buf, pages, npages.value, doff.value, dcount.value)


def decode_chunks(names, chunks, stream=None, arrow_types=None) -> ColumnarBatch:
    lib = capi.load()
    arr = (capi.sb_column_chunk * len(chunks))()
    for i, ch in enumerate(chunks):
        arr[i] = ch.c()
    h = C.c_void_p()
    capi.check(lib.sb_scan_decode(arr, len(chunks), _h(stream), C.byref(h)))
    return ColumnarBatch(h, list(names), arrow_types)


class ParquetScanExec(SparkPlan):
    """Scan of a Parquet file: `batches()` yields one HBM-resident ColumnarBatch per row group; executeColumnar concatenates
    them (for plans that want the partition as one batch).  Pages must be uncompressed."""

    def __init__(self, path, columns=None):
        self.path = path
        self.file = pq.ParquetFile(path)
        self.meta = self.file.metadata
        schema = self.file.schema_arrow
        self.columns = list(columns) if columns is not None else list(schema.names)
        self.fields = [schema.field(c) for c in self.columns]
        self._f = open(path, "rb")
        self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)
        self._bytes = np.frombuffer(self._mm, dtype=np.uint8)

    def row_group_chunks(self, rg: int):
        g = self.meta.row_group(rg)
        by_name = {g.column(i).path_in_schema: g.column(i) for i in range(g.num_columns)}
        out = []
        for f in self.fields:
            cm = by_name[f.name]
            if cm.compression != "UNCOMPRESSED":
                raise capi.SparkB200Error(5, "column %s is %s-compressed; the GPU scan reads uncompressed pages" % (f.name, cm.compression))
            if cm.physical_type not in _PHYS:
                raise capi.SparkB200Error(5, "column %s: physical type %s is not supported by the GPU scan" % (f.name, cm.physical_type))
            start = cm.dictionary_page_offset if cm.has_dictionary_page and cm.dictionary_page_offset else cm.data_page_offset
            start = min(start, cm.data_page_offset)
            buf = self._bytes[start:start + cm.total_compressed_size]
            t, scale = _sb_type(f.type)
            max_def = self.file.schema.column(self.file.schema.names.index(f.name)).max_definition_level
            out.append(EncodedColumnChunk.from_file_bytes(buf, t, scale, _PHYS[cm.physical_type], max_def))
        return out

    def batches(self, stream=None):
        for rg in range(self.meta.num_row_groups):
            yield decode_chunks(self.columns, self.row_group_chunks(rg), stream, [f.type for f in self.fields])

    def executeColumnar(self, stream=None):
        bs = list(self.batches(stream))
        if len(bs) == 1:
            return bs[0]
        try:
            return ColumnarBatch.concat(bs, stream)
        finally:
            for b in bs:
                b.close()


def encode_column(batch: ColumnarBatch, name: str, dictionary: bool, page_rows: int = 1 << 20, stream=None, pinned=False) -> EncodedColumnChunk:
    """Write-side twin (tests, bench.py): encodes one NULL-free fixed-width column the way a Parquet writer would -- a sorted
    dictionary + pages of bit-packed indices (RLE_DICTIONARY), or PLAIN pages -- entirely on the GPU (distinct values through the
    hash aggregate, order through the radix sort, indices + bit packing in sb_scan_encode) and brings the chunk to the host."""
    from .execution import HashAggregateExec, LocalTableScanExec, SortExec
    from .columnar import PinnedArray
    lib = capi.load()
    col = batch.column_index(name)
    d = batch.column_desc(col)
    n = batch.num_rows
    dict_batch = None
    if dictionary:
        one = batch.select([name])
        try:
            dict_batch = SortExec([(name, True, True)], HashAggregateExec([name], [], LocalTableScanExec(one), mode="complete")).executeColumnar(stream)
        finally:
            one.close()
    cap = max(1, (n + page_rows - 1) // page_rows)
    pages = (capi.sb_page * cap)()
    npages, doff, dcount = C.c_int32(), C.c_int64(), C.c_int32()
    h = C.c_void_p()
    try:
        capi.check(lib.sb_scan_encode(batch.handle, col, dict_batch.handle if dict_batch else None, page_rows, _h(stream), C.byref(h), pages, cap,
                                      C.byref(npages), C.byref(doff), C.byref(dcount)))
    finally:
        if dict_batch:
            dict_batch.close()
    chunk = ColumnarBatch(h, ["bytes"], [pa.int8()])
    try:
        nbytes = chunk.num_rows
        if pinned:
            keep = PinnedArray(max(nbytes, 1), np.uint8)
            host = keep.array[:nbytes]
        else:
            keep = None
            host = np.empty(nbytes, np.uint8)
        capi.check(lib.sb_table_export_host(chunk.handle, 0, host.ctypes.data if nbytes else None, None, None, None, _h(stream)))
    finally:
        chunk.close()
    width = capi.TYPE_WIDTH[d.type]
    phys = (capi.SB_PHYS["DOUBLE"] if d.type == capi.SB_FLOAT64 else capi.SB_PHYS["FLOAT"] if d.type == capi.SB_FLOAT32
            else capi.SB_PHYS["INT64"] if width == 8 else capi.SB_PHYS["INT32"])
    ech = EncodedColumnChunk(d.type, d.scale, phys, host, pages, npages.value, doff.value, dcount.value)
    ech._keep = keep
    return ech

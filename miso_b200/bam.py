"""BAM (+ BAI) reading without pysam: the alignment input of ``compute_gene_psi``.

The reference opens its alignments with ``pysam.Samfile(bam, "rb")`` and asks for the reads of a
gene with ``bamfile.fetch(chrom, start, end)`` (``misopy/sam_utils.py:143-181``); per read it uses
``qname``, ``flag``, ``pos`` (0-based), ``cigar`` and ``rlen`` (``sam_utils.py:186-456``).  ``pysam``
is not in this image, so this module restates the two published formats it wraps (SAM/BAM
specification, sections 4.1 "BGZF", 4.2 "BAM", 5.2 "BAI"):

  * BGZF: a series of gzip members of at most 64 KiB, each with a ``BC`` extra field holding the
    compressed block size; a *virtual offset* is ``(file offset of the block) << 16 | offset in
    the block's data``;
  * BAM: magic, header text, reference names and lengths, then alignment records
    (``block_size, refID, pos, l_read_name, mapq, bin, n_cigar_op, flag, l_seq, next_refID,
    next_pos, tlen, read_name, cigar[], seq, qual, tags``);
  * BAI: per reference the bins (lists of chunks of virtual offsets) of the UCSC binning scheme
    and a 16 kb linear index.

``BamFile.fetch`` yields the reads whose alignment overlaps ``[start, end)`` on a reference, in
file order -- with the index when ``<bam>.bai`` (or ``<stem>.bai``) exists, by a scan of the file
otherwise -- as ``run_miso.SamRead`` tuples, the record the SAM text loader produces, so everything
downstream (pairing, strand rules, read-length filter) is shared.
"""
import os
import struct
import zlib
from collections import namedtuple

SamRead = namedtuple("SamRead", "qname flag rname pos cigar rlen aend")     # pos, aend 0-based, aend exclusive

_CIGAR_OPS = "MIDNSHP=X"
_CONSUMES_REF = (True, False, True, True, False, False, False, True, True)


class BamError(ValueError):
    pass


class _Bgzf:
    """Random access to the uncompressed stream of a BGZF file by virtual offset."""

    def __init__(self, path):
        self.f = open(path, "rb")
        self.block_start = -1          # file offset of the block held in self.data
        self.block_len = 0             # its compressed size
        self.data = b""
        self.off = 0                   # position in self.data

    def close(self):
        self.f.close()

    def _load(self, start):
        self.f.seek(start)
        head = self.f.read(12)
        if len(head) == 0:
            self.block_start, self.block_len, self.data, self.off = start, 0, b"", 0
            return False
        if len(head) < 12 or head[:4] != b"\x1f\x8b\x08\x04":
            raise BamError("not a BGZF block at file offset %d" % start)
        xlen = struct.unpack("<H", head[10:12])[0]
        extra = self.f.read(xlen)
        bsize, i = None, 0
        while i + 4 <= len(extra):
            si1, si2, slen = extra[i], extra[i + 1], struct.unpack("<H", extra[i + 2:i + 4])[0]
            if si1 == 66 and si2 == 67 and slen == 2:
                bsize = struct.unpack("<H", extra[i + 4:i + 6])[0] + 1
            i += 4 + slen
        if bsize is None:
            raise BamError("BGZF block without a BC field at file offset %d" % start)
        cdata = self.f.read(bsize - 12 - xlen - 8)
        tail = self.f.read(8)
        if len(tail) < 8:
            raise BamError("truncated BGZF block at file offset %d" % start)
        data = zlib.decompress(cdata, -15)
        crc, isize = struct.unpack("<II", tail)
        if len(data) != isize or (zlib.crc32(data) & 0xffffffff) != crc:
            raise BamError("corrupt BGZF block at file offset %d" % start)
        self.block_start, self.block_len, self.data, self.off = start, bsize, data, 0
        return True

    def seek(self, voffset):
        start, within = voffset >> 16, voffset & 0xffff
        if start != self.block_start:
            self._load(start)
        self.off = within

    def tell(self):
        if self.off >= len(self.data) and self.block_len:       # the position after a block is the next one's start
            return (self.block_start + self.block_len) << 16
        return (self.block_start << 16) | self.off

    def read(self, n):
        out = []
        while n > 0:
            if self.off >= len(self.data):
                if self.block_start >= 0 and self.block_len == 0:
                    break                                        # end of file
                if not self._load(self.block_start + self.block_len if self.block_start >= 0 else 0):
                    break
                continue
            chunk = self.data[self.off:self.off + n]
            self.off += len(chunk)
            n -= len(chunk)
            out.append(chunk)
        return b"".join(out)


def reg2bins(beg, end):
    """Bins that may hold alignments overlapping [beg, end) (SAM specification, section 5.3)."""
    end -= 1
    bins = [0]
    for shift, first in ((26, 1), (23, 9), (20, 73), (17, 585), (14, 4681)):
        bins.extend(range(first + (beg >> shift), first + (end >> shift) + 1))
    return bins


class BamFile:
    """``pysam.Samfile(path, "rb")`` as far as ``misopy`` uses it: ``references``, ``lengths``,
    ``header_text``, ``fetch(chrom, start, end)`` and iteration over all records."""

    def __init__(self, path, index=None):
        self.path = path
        self._z = _Bgzf(path)
        z = self._z
        z.seek(0)
        if z.read(4) != b"BAM\x01":
            raise BamError("%s is not a BAM file" % path)
        l_text = struct.unpack("<i", z.read(4))[0]
        self.header_text = z.read(l_text).split(b"\x00", 1)[0].decode("ascii", "replace")
        n_ref = struct.unpack("<i", z.read(4))[0]
        self.references, self.lengths = [], []
        for _ in range(n_ref):
            l_name = struct.unpack("<i", z.read(4))[0]
            self.references.append(z.read(l_name).rstrip(b"\x00").decode("ascii"))
            self.lengths.append(struct.unpack("<i", z.read(4))[0])
        self._first = z.tell()
        self._tid = {name: i for i, name in enumerate(self.references)}
        self._index = None
        for cand in ([index] if index else [path + ".bai", os.path.splitext(path)[0] + ".bai"]):
            if cand and os.path.isfile(cand):
                self._index = _load_bai(cand, n_ref)
                break

    def close(self):
        self._z.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    @property
    def has_index(self):
        return self._index is not None

    def _record(self):
        z = self._z
        head = z.read(4)
        if len(head) < 4:
            return None
        block_size = struct.unpack("<i", head)[0]
        rec = z.read(block_size)
        if len(rec) < block_size or block_size < 32:
            raise BamError("truncated alignment record in %s" % self.path)
        ref_id, pos, l_read_name, _mapq, _bin, n_cigar, flag, l_seq = struct.unpack("<iiBBHHHi", rec[:20])
        name = rec[32:32 + l_read_name - 1].decode("ascii", "replace")
        o = 32 + l_read_name
        cigar, span = None, 0
        if n_cigar:
            ops = struct.unpack("<%dI" % n_cigar, rec[o:o + 4 * n_cigar])
            parts = []
            for v in ops:
                op, ln = v & 0xf, v >> 4
                if op >= len(_CIGAR_OPS):
                    raise BamError("unknown CIGAR operation %d in %s" % (op, self.path))
                parts.append("%d%s" % (ln, _CIGAR_OPS[op]))
                if _CONSUMES_REF[op]:
                    span += ln
            cigar = "".join(parts)
        rname = self.references[ref_id] if 0 <= ref_id < len(self.references) else "*"
        return ref_id, SamRead(name, flag, rname, pos, cigar, l_seq, pos + span)

    def __iter__(self):
        self._z.seek(self._first)
        while True:
            r = self._record()
            if r is None:
                return
            yield r[1]

    def fetch(self, chrom, start, end):
        """Reads of ``chrom`` whose alignment overlaps [start, end) (0-based, end exclusive), file order.
        Raises ``ValueError`` for an unknown reference name, as pysam does."""
        if chrom not in self._tid:
            raise ValueError("invalid reference `%s`" % chrom)
        tid = self._tid[chrom]
        start, end = max(0, int(start)), int(end)
        if end <= start:
            return []
        out = []

        def take(read):
            # a read without CIGAR covers one base (pysam / htslib: bam_endpos)
            aend = read.aend if read.aend > read.pos else read.pos + 1
            return read.pos < end and aend > start

        if self._index is None:
            self._z.seek(self._first)
            while True:
                r = self._record()
                if r is None:
                    break
                if r[0] == tid and take(r[1]):
                    out.append(r[1])
            return out
        bins, linear = self._index[tid]
        min_off = 0
        if linear:
            w = start >> 14
            min_off = linear[w] if w < len(linear) else linear[-1]
        chunks = sorted(c for b in reg2bins(start, end) for c in bins.get(b, ()) if c[1] > min_off)
        merged = []
        for beg, fin in chunks:
            if merged and beg <= merged[-1][1]:
                merged[-1][1] = max(merged[-1][1], fin)
            else:
                merged.append([beg, fin])
        for beg, fin in merged:
            self._z.seek(beg)
            while self._z.tell() < fin:
                r = self._record()
                if r is None:
                    break
                if r[0] != tid or r[1].pos >= end:
                    break                                       # sorted file: nothing further in this chunk
                if take(r[1]):
                    out.append(r[1])
        return out


def _load_bai(path, n_ref_bam):
    with open(path, "rb") as f:
        b = f.read()
    if b[:4] != b"BAI\x01":
        raise BamError("%s is not a BAI index" % path)
    n_ref = struct.unpack_from("<i", b, 4)[0]
    o = 8
    refs = []
    for _ in range(n_ref):
        n_bin = struct.unpack_from("<i", b, o)[0]
        o += 4
        bins = {}
        for _ in range(n_bin):
            bin_id, n_chunk = struct.unpack_from("<Ii", b, o)
            o += 8
            chunks = [struct.unpack_from("<QQ", b, o + 16 * i) for i in range(n_chunk)]
            o += 16 * n_chunk
            if bin_id != 37450:                                  # pseudo-bin with mapped / unmapped counts
                bins[bin_id] = chunks
        n_intv = struct.unpack_from("<i", b, o)[0]
        o += 4
        linear = list(struct.unpack_from("<%dQ" % n_intv, b, o)) if n_intv else []
        o += 8 * n_intv
        refs.append((bins, linear))
    while len(refs) < n_ref_bam:
        refs.append(({}, []))
    return refs


# ---- writing (tests and fixtures): SAM-like records -> BGZF-compressed BAM -------------------
def _reg2bin(beg, end):
    end -= 1
    for shift, first in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
        if beg >> shift == end >> shift:
            return first + (beg >> shift)
    return 0


def write_bam(path, references, reads, header_text="", block_bytes=0xff00, index_path=None):
    """A coordinate-ordered list of ``SamRead`` (sequence and qualities are filled with ``N`` / 0xff)
    as a BAM file, with its BAI index when ``index_path`` is given.  ``references``: [(name, length)].
    Used by the tests; the product only reads."""
    tid = {name: i for i, (name, _) in enumerate(references)}
    raw = [b"BAM\x01", struct.pack("<i", len(header_text)), header_text.encode("ascii"),
           struct.pack("<i", len(references))]
    for name, length in references:
        raw.append(struct.pack("<i", len(name) + 1) + name.encode("ascii") + b"\x00" + struct.pack("<i", length))
    at = sum(len(x) for x in raw)
    spans = []                                   # per read: (tid, pos, end, bin, uncompressed begin, end)
    for r in reads:
        ops = []
        if r.cigar:
            num = ""
            for ch in r.cigar:
                if ch.isdigit():
                    num += ch
                else:
                    ops.append((int(num) << 4) | _CIGAR_OPS.index(ch))
                    num = ""
        name = r.qname.encode("ascii") + b"\x00"
        span = max(1, r.aend - r.pos)
        rbin = _reg2bin(r.pos, r.pos + span)
        body = struct.pack("<iiBBHHHiiii", tid.get(r.rname, -1), r.pos, len(name), 255, rbin,
                           len(ops), r.flag, r.rlen, -1, -1, 0)
        body += name + struct.pack("<%dI" % len(ops), *ops) + b"\xff" * ((r.rlen + 1) // 2) + b"\xff" * r.rlen
        raw.append(struct.pack("<i", len(body)) + body)
        spans.append((tid.get(r.rname, -1), r.pos, r.pos + span, rbin, at, at + 4 + len(body)))
        at += 4 + len(body)
    data = b"".join(raw)
    block_at = []                                # file offset of every block
    with open(path, "wb") as f:
        for i in list(range(0, len(data), block_bytes)) + [None]:
            chunk = b"" if i is None else data[i:i + block_bytes]          # the last, empty block is the EOF marker
            c = zlib.compressobj(6, zlib.DEFLATED, -15)
            comp = c.compress(chunk) + c.flush()
            block_at.append(f.tell())
            f.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" +
                    struct.pack("<H", len(comp) + 25) + comp +
                    struct.pack("<II", zlib.crc32(chunk) & 0xffffffff, len(chunk)))
    if index_path is None:
        return

    def voff(u):
        return (block_at[u // block_bytes] << 16) | (u % block_bytes)

    out = [b"BAI\x01", struct.pack("<i", len(references))]
    for t in range(len(references)):
        bins, linear = {}, []
        for rt, pos, end, rbin, u0, u1 in spans:
            if rt != t:
                continue
            chunks = bins.setdefault(rbin, [])
            if chunks and chunks[-1][1] == voff(u0):
                chunks[-1][1] = voff(u1)
            else:
                chunks.append([voff(u0), voff(u1)])
            for w in range(pos >> 14, ((end - 1) >> 14) + 1):
                while len(linear) <= w:
                    linear.append(0)
                if linear[w] == 0:
                    linear[w] = voff(u0)
        for w in range(1, len(linear)):          # windows without a read inherit the previous offset
            if linear[w] == 0:
                linear[w] = linear[w - 1]
        out.append(struct.pack("<i", len(bins)))
        for b, chunks in sorted(bins.items()):
            out.append(struct.pack("<Ii", b, len(chunks)) + b"".join(struct.pack("<QQ", c0, c1) for c0, c1 in chunks))
        out.append(struct.pack("<i", len(linear)) + struct.pack("<%dQ" % len(linear), *linear))
    with open(index_path, "wb") as f:
        f.write(b"".join(out))

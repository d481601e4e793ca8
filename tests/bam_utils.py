"""Test infrastructure: a minimal BAM writer (SAM/BAM specification 4.1-4.2: BGZF blocks of <= 64 KiB + the empty end-of-file block).
No htslib / pysam / samtools in this image, so the fixtures for dropest_b200/host/BamIngest are made here."""
import struct
import zlib


def _bgzf_block(data: bytes, level: int = 6) -> bytes:
    comp = zlib.compressobj(level, zlib.DEFLATED, -15)
    body = comp.compress(data) + comp.flush()
    bsize = len(body) + 25  # header 18 + trailer 8 - 1
    assert bsize < 65536
    return (b"\x1f\x8b\x08\x04" + b"\x00" * 4 + b"\x00\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, bsize) + body +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def _tag(name: str, value) -> bytes:
    kind, v = value
    if kind == "Z":
        return name.encode() + b"Z" + v.encode() + b"\x00"
    if kind == "A":
        return name.encode() + b"A" + v.encode()[:1]
    if kind == "i":
        return name.encode() + b"i" + struct.pack("<i", v)
    if kind == "C":
        return name.encode() + b"C" + struct.pack("<B", v)
    if kind == "B":  # array of uint16
        return name.encode() + b"BS" + struct.pack("<I", len(v)) + b"".join(struct.pack("<H", x) for x in v)
    raise ValueError(kind)


CIGAR_OPS = "MIDNSHP=X"


def alignment(name: str, ref_id: int, pos: int, flag: int, tags, seq_len: int = 8, cigar_ops=None) -> bytes:
    """tags: list of (tag, (type, value)); sequence / qualities are filler (the ingest does not read them); cigar_ops: list of
    (op letter, length), default one match of seq_len"""
    rn = name.encode() + b"\x00"
    if cigar_ops is None:
        cigar_ops = [("M", seq_len)]
    cigar = b"".join(struct.pack("<I", (length << 4) | CIGAR_OPS.index(op)) for op, length in cigar_ops)
    seq = b"\x11" * ((seq_len + 1) // 2)
    qual = b"\x1e" * seq_len
    tag_bytes = b"".join(_tag(t, v) for t, v in tags)
    core = struct.pack("<iiBBHHHIiii", ref_id, pos, len(rn), 30, 4680, len(cigar_ops), flag, seq_len, -1, -1, 0)
    rec = core + rn + cigar + seq + qual + tag_bytes
    return struct.pack("<I", len(rec)) + rec


def write_bam(path: str, references, alignments, block_bytes: int = 60000, header_text: str = "@HD\tVN:1.6\n", level: int = 6):
    """references: list of (name, length); alignments: iterable of bytes from alignment()"""
    head = b"BAM\x01" + struct.pack("<I", len(header_text)) + header_text.encode() + struct.pack("<I", len(references))
    for name, length in references:
        head += struct.pack("<I", len(name) + 1) + name.encode() + b"\x00" + struct.pack("<I", length)
    payload = head + b"".join(alignments)
    with open(path, "wb") as f:
        for off in range(0, len(payload), block_bytes):   # records freely straddle blocks, as in real files
            f.write(_bgzf_block(payload[off:off + block_bytes], level))
        f.write(_bgzf_block(b""))


def read_bam(path: str):
    """(header text, [(reference name, length)], [(read name, ref id, pos, flag, [TAG:TYPE:VALUE, ...])]) of a BAM file, and checks the BGZF
    framing on the way: every block is a gzip member with the BC subfield, <= 64 KiB, CRC and size match, the file ends with the empty block."""
    raw = open(path, "rb").read()
    off, payload, last_isize = 0, [], None
    while off < len(raw):
        assert raw[off:off + 4] == b"\x1f\x8b\x08\x04" and raw[off + 12:off + 16] == b"BC\x02\x00", "not a BGZF block"
        xlen = struct.unpack_from("<H", raw, off + 10)[0]
        bsize = struct.unpack_from("<H", raw, off + 16)[0] + 1
        assert xlen == 6 and bsize <= 65536 and off + bsize <= len(raw)
        crc, isize = struct.unpack_from("<II", raw, off + bsize - 8)
        data = zlib.decompress(raw[off + 18:off + bsize - 8], -15)
        assert len(data) == isize and (zlib.crc32(data) & 0xFFFFFFFF) == crc and isize <= 65536
        payload.append(data)
        last_isize = isize
        off += bsize
    assert last_isize == 0, "no end-of-file block"
    buf = b"".join(payload)
    assert buf[:4] == b"BAM\x01"
    l_text = struct.unpack_from("<I", buf, 4)[0]
    text = buf[8:8 + l_text].decode()
    p = 8 + l_text
    n_ref = struct.unpack_from("<I", buf, p)[0]
    p += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<I", buf, p)[0]
        name = buf[p + 4:p + 4 + l_name - 1].decode()
        refs.append((name, struct.unpack_from("<I", buf, p + 4 + l_name)[0]))
        p += 8 + l_name
    records = []
    while p < len(buf):
        block_size = struct.unpack_from("<I", buf, p)[0]
        rec = buf[p + 4:p + 4 + block_size]
        assert len(rec) == block_size
        p += 4 + block_size
        ref_id, pos, l_rn, _mapq, _bin, n_cigar, flag, l_seq = struct.unpack_from("<iiBBHHHI", rec, 0)
        name = rec[32:32 + l_rn - 1].decode()
        q = 32 + l_rn + 4 * n_cigar + (l_seq + 1) // 2 + l_seq
        tags = []
        while q < len(rec):
            tag, t = rec[q:q + 2].decode(), chr(rec[q + 2])
            q += 3
            if t in "ZH":
                e = rec.index(b"\x00", q)
                tags.append(f"{tag}:{t}:{rec[q:e].decode()}")
                q = e + 1
            elif t == "A":
                tags.append(f"{tag}:A:{chr(rec[q])}")
                q += 1
            elif t == "B":
                st, cnt = chr(rec[q]), struct.unpack_from("<I", rec, q + 1)[0]
                es = {"c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}[st]
                fmt = {"c": "b", "C": "B", "s": "h", "S": "H", "i": "i", "I": "I", "f": "f"}[st]
                vals = struct.unpack_from("<" + fmt * cnt, rec, q + 5)
                tags.append(f"{tag}:B:{st}," + ",".join(str(v) for v in vals))
                q += 5 + cnt * es
            else:
                es, fmt = {"c": (1, "b"), "C": (1, "B"), "s": (2, "h"), "S": (2, "H"), "i": (4, "i"), "I": (4, "I"), "f": (4, "f")}[t]
                tags.append(f"{tag}:{t}:{struct.unpack_from('<' + fmt, rec, q)[0]}")
                q += es
        records.append((name, ref_id, pos, flag, tags))
    return text, refs, records

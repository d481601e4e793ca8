"""Test infrastructure: a minimal BAM writer (SAM/BAM specification 4.1-4.2: BGZF blocks of <= 64 KiB + the empty end-of-file block).
No htslib / pysam / samtools in this image, so the fixtures for dropest_b200/host/BamIngest are made here."""
import struct
import zlib


def _bgzf_block(data: bytes) -> bytes:
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
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


def write_bam(path: str, references, alignments, block_bytes: int = 60000, header_text: str = "@HD\tVN:1.6\n"):
    """references: list of (name, length); alignments: iterable of bytes from alignment()"""
    head = b"BAM\x01" + struct.pack("<I", len(header_text)) + header_text.encode() + struct.pack("<I", len(references))
    for name, length in references:
        head += struct.pack("<I", len(name) + 1) + name.encode() + b"\x00" + struct.pack("<I", length)
    payload = head + b"".join(alignments)
    with open(path, "wb") as f:
        for off in range(0, len(payload), block_bytes):   # records freely straddle blocks, as in real files
            f.write(_bgzf_block(payload[off:off + block_bytes]))
        f.write(_bgzf_block(b""))

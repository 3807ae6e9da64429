"""`.ksplat` container (the in-memory SplatBuffer format, src/loaders/SplatBuffer.js): header parsing and a writer.

Only the byte layout lives here (host side, no arithmetic on the hot path).  Decoding the records into the arrays the
rasteriser and the sorter consume happens ON THE GPU (csrc/ksplat_kernels.cuh, C ABI gs_upload_ksplat); the NumPy decoder
used to check it lives with the test infrastructure, not in this package.

Layout (all little-endian):
  [header 4096 B][maxSectionCount x section header 1024 B][section 0][section 1]...
  section = [partiallyFilledBucketLengths u32 x P][bucket centres f32x3 x B][records bytesPerSplat x maxSplatCount]
  record  = centre (3xf32 | 3xu16 bucket-relative) | scale 3x(f32|f16) | rotation w,x,y,z 4x(f32|f16) | rgba 4xu8 | SH (f32|f16|u8)
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

HEADER_BYTES = 4096            # SplatBuffer.HeaderSizeBytes        (SplatBuffer.js:167)
SECTION_HEADER_BYTES = 1024    # SplatBuffer.SectionHeaderSizeBytes (:168)
BUCKET_STORAGE_BYTES = 12      # :170
BUCKET_BLOCK_SIZE = 5.0        # :173
BUCKET_SIZE = 256              # :174
SH8_RANGE = 3.0                # Constants.SphericalHarmonics8BitCompressionRange (src/Constants.js:10)

# SplatBuffer.CompressionLevels (:108-163)
LEVELS = {
    0: dict(center=12, scale=12, rot=16, color=4, sh_comp=4, scale_range=1),
    1: dict(center=6, scale=6, rot=8, color=4, sh_comp=2, scale_range=32767),
    2: dict(center=6, scale=6, rot=8, color=4, sh_comp=1, scale_range=32767),
}
SH_COMPONENTS = {0: 0, 1: 9, 2: 24}


def bytes_per_splat(level: int, sh_degree: int) -> int:
    """SplatBuffer.calculateComponentStorage (:992-1010)."""
    L = LEVELS[level]
    return L["center"] + L["scale"] + L["rot"] + L["color"] + L["sh_comp"] * SH_COMPONENTS[sh_degree]


@dataclass
class Section:
    splat_count: int
    max_splat_count: int
    bucket_size: int
    bucket_count: int
    bucket_block_size: float
    bucket_storage_bytes: int
    compression_scale_range: int
    full_bucket_count: int
    partially_filled_bucket_count: int
    sh_degree: int
    bytes_per_splat: int
    base: int                 # file offset of the section (partial-bucket lengths)
    buckets_base: int         # file offset of the bucket centres
    data_base: int            # file offset of the first record
    splat_count_offset: int   # global index of the section's first splat


@dataclass
class Header:
    version: tuple[int, int]
    max_section_count: int
    section_count: int
    max_splat_count: int
    splat_count: int
    compression_level: int
    scene_center: tuple[float, float, float]
    min_sh: float
    max_sh: float
    sections: list[Section] = field(default_factory=list)


def parse(data: bytes | np.ndarray) -> Header:
    """SplatBuffer.parseHeader (:819-848) + parseSectionHeaders (:877-941); counts as with secLoadedCountsToMax."""
    buf = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data.view(np.uint8)
    if buf.size < HEADER_BYTES:
        raise ValueError("not a .ksplat: shorter than the 4096-byte header")
    u8, u16, u32, f32 = buf[:HEADER_BYTES], buf[:HEADER_BYTES].view(np.uint16), buf[:HEADER_BYTES].view(np.uint32), buf[:HEADER_BYTES].view(np.float32)
    level = int(u16[10])
    if level not in LEVELS:
        raise ValueError(f"unknown compression level {level}")
    h = Header((int(u8[0]), int(u8[1])), int(u32[1]), int(u32[2]), int(u32[3]), int(u32[4]), level,
               (float(f32[6]), float(f32[7]), float(f32[8])), float(f32[9]) or -SH8_RANGE / 2, float(f32[10]) or SH8_RANGE / 2)
    if (h.version[0], h.version[1]) < (0, 1):  # KSplatLoader.checkVersion (KSplatLoader.js:8-20)
        raise ValueError(f"unsupported .ksplat version {h.version}")
    base = HEADER_BYTES + h.max_section_count * SECTION_HEADER_BYTES
    offset = 0
    for i in range(h.max_section_count):
        sh = buf[HEADER_BYTES + i * SECTION_HEADER_BYTES: HEADER_BYTES + (i + 1) * SECTION_HEADER_BYTES]
        s16, s32, sf = sh.view(np.uint16), sh.view(np.uint32), sh.view(np.float32)
        max_count, bucket_size, bucket_count = int(s32[1]), int(s32[2]), int(s32[3])
        storage = int(s16[10])
        scale_range = int(s32[6]) or LEVELS[level]["scale_range"]
        full, partial = int(s32[8]), int(s32[9])
        deg = int(s16[20])
        bps = bytes_per_splat(level, deg)
        meta = partial * 4
        buckets_bytes = storage * bucket_count + meta
        h.sections.append(Section(max_count, max_count, bucket_size, bucket_count, float(sf[4]), storage, scale_range, full, partial, deg, bps,
                                  base, base + meta, base + buckets_bytes, offset))
        base += bps * max_count + buckets_bytes
        offset += max_count
    return h


def to_half_three(a: np.ndarray) -> np.ndarray:
    """THREE.DataUtils.toHalfFloat (three r160): the table-driven conversion base[e] + (mantissa >> shift[e]) -- it TRUNCATES the
    mantissa (no rounding) after clamping to +-65504.  Returns the f16 bit patterns.  (three is not vendored: restated from its
    published algorithm, parity unpinned.)"""
    f = np.clip(np.asarray(a, np.float32), -65504.0, 65504.0).astype(np.float32).view(np.uint32)
    sign = (f >> 16) & 0x8000
    mant = f & 0x007FFFFF
    ex = ((f >> 23) & 0xFF).astype(np.int64) - 127
    base = np.zeros(f.shape, np.uint32)
    shift = np.full(f.shape, 24, np.uint32)
    m = (ex >= -27) & (ex < -14)
    base[m] = (0x0400 >> (-ex[m] - 14)).astype(np.uint32); shift[m] = (-ex[m] - 1).astype(np.uint32)
    m = (ex >= -14) & (ex <= 15)
    base[m] = ((ex[m] + 15) << 10).astype(np.uint32); shift[m] = 13
    m = (ex > 15) & (ex < 128)
    base[m] = 0x7C00; shift[m] = 24
    m = ex >= 128
    base[m] = 0x7C00; shift[m] = 13
    return ((base | sign) + (mant >> shift)).astype(np.uint16)


_half_bits = to_half_three


def compute_buckets(centers: np.ndarray, block_size: float, bucket_size: int):
    """SplatBuffer.computeBucketsForUncompressedSplatArray (:1328-1399): (ordered splat rows, bucket centres, #full, partial lengths).
    Full buckets in completion order, then the partially filled ones in ascending bucket id (JS integer-key order)."""
    c = centers.astype(np.float64)
    mn = c.min(0)
    dims = c.max(0) - mn
    yb, zb = int(np.ceil(dims[1] / block_size)), int(np.ceil(dims[2] / block_size))
    blk = np.floor((c - mn) / block_size).astype(np.int64)
    ids = blk[:, 0] * (yb * zb) + blk[:, 1] * zb + blk[:, 2]
    centre_of = blk * block_size + mn + block_size / 2.0
    full, open_ = [], {}
    for i, b in enumerate(ids.tolist()):
        cur = open_.get(b)
        if cur is None:
            cur = open_[b] = ([], centre_of[i])
        cur[0].append(i)
        if len(cur[0]) >= bucket_size:
            full.append(cur)
            open_[b] = None
    partial = [open_[b] for b in sorted(open_) if open_[b] is not None]
    buckets = full + partial
    rows = np.concatenate([np.asarray(b[0], np.int64) for b in buckets]) if buckets else np.zeros(0, np.int64)
    centres = np.stack([b[1] for b in buckets]).astype(np.float32) if buckets else np.zeros((0, 3), np.float32)
    return rows, centres, len(full), [len(b[0]) for b in partial]


def write(centers, scales, rotations_xyzw, colors, sh=None, sh_degree=0, *, compression_level=0, minimum_alpha=1,
          block_size=BUCKET_BLOCK_SIZE, bucket_size=BUCKET_SIZE, scene_center=(0.0, 0.0, 0.0)) -> bytes:
    """One-section .ksplat like SplatBuffer.generateFromUncompressedSplatArrays (:1177-1326).

    `sh`: [n, ncoef, 3] in the GPU-side order (coefficient-major RGB triples); written channel-major per band, the file's order."""
    level = compression_level
    keep = colors[:, 3].astype(np.int64) >= minimum_alpha
    centers, scales, rot, colors = centers[keep].astype(np.float32), scales[keep].astype(np.float32), rotations_xyzw[keep].astype(np.float32), colors[keep]
    n = centers.shape[0]
    ncomp = SH_COMPONENTS[sh_degree]
    shf = None
    lo, hi = -SH8_RANGE / 2, SH8_RANGE / 2
    if ncomp:
        t = sh[keep].astype(np.float32)                                   # [n, ncoef, 3] -> file order
        parts = [t[:, :3].transpose(0, 2, 1).reshape(n, 9)]               # band 1: [R: c1..c3][G][B]
        if sh_degree >= 2:
            parts.append(t[:, 3:8].transpose(0, 2, 1).reshape(n, 15))    # band 2: [R: c4..c8][G][B]
        shf = np.concatenate(parts, 1)
        nz = shf[shf != 0]
        lo = float(shf.min()) if shf.size and shf.min() != 0 else (float(nz.min()) if nz.size else lo)   # the writer's "falsy" quirk (:1189-1199)
        hi = float(shf.max()) if shf.size and shf.max() != 0 else (float(nz.max()) if nz.size else hi)
    bps = bytes_per_splat(level, sh_degree)
    L = LEVELS[level]
    q = rot / np.linalg.norm(rot, axis=1, keepdims=True)
    wxyz = np.stack([q[:, 3], q[:, 0], q[:, 1], q[:, 2]], 1)             # file order = PLY rot_0..3 = w,x,y,z

    rows = np.arange(n, dtype=np.int64)
    bucket_centres = np.zeros((0, 3), np.float32)
    full_count, partial_lengths = 0, []
    if level >= 1:
        rows, bucket_centres, full_count, partial_lengths = compute_buckets(centers, block_size, bucket_size)
    rec = np.zeros((n, bps), np.uint8)
    src = rows
    if level == 0:
        rec[:, 0:12] = centers[src].view(np.uint8).reshape(n, 12)
        rec[:, 12:24] = scales[src].view(np.uint8).reshape(n, 12)
        rec[:, 24:40] = wxyz[src].astype(np.float32).view(np.uint8).reshape(n, 16)
        rec[:, 40:44] = colors[src]
        if ncomp:
            rec[:, 44:44 + 4 * ncomp] = shf[src].astype(np.float32).view(np.uint8).reshape(n, 4 * ncomp)
    else:
        scale_range = L["scale_range"]
        factor = scale_range / (block_size * 0.5)
        lengths = [bucket_size] * full_count + partial_lengths
        bidx = np.repeat(np.arange(len(lengths)), lengths)
        delta = centers[src].astype(np.float64) - bucket_centres[bidx].astype(np.float64)
        cu = np.clip(np.floor(delta * factor + 0.5) + scale_range, 0, scale_range * 2 + 1).astype(np.uint16)   # Math.round, clamp (:1068-1072)
        rec[:, 0:6] = cu.view(np.uint8).reshape(n, 6)
        rec[:, 6:12] = _half_bits(scales[src]).view(np.uint8).reshape(n, 6)
        rec[:, 12:20] = _half_bits(wxyz[src]).view(np.uint8).reshape(n, 8)
        rec[:, 20:24] = colors[src]
        if ncomp:
            if level == 1:
                rec[:, 24:24 + 2 * ncomp] = _half_bits(shf[src]).view(np.uint8).reshape(n, 2 * ncomp)
            else:  # toUint8 (:22-26)
                v = np.clip(shf[src].astype(np.float64), lo, hi)
                rec[:, 24:24 + ncomp] = np.clip(np.floor((v - lo) / (hi - lo) * 255), 0, 255).astype(np.uint8)

    meta = np.asarray(partial_lengths, np.uint32).tobytes() if level >= 1 else b""
    bucket_bytes = bucket_centres.astype(np.float32).tobytes() if level >= 1 else b""
    section = meta + bucket_bytes + rec.tobytes()
    header = np.zeros(HEADER_BYTES, np.uint8)
    header[0], header[1] = 0, 1
    h32, h16, hf = header.view(np.uint32), header.view(np.uint16), header.view(np.float32)
    h32[1], h32[2], h32[3], h32[4] = 1, 1, n, n
    h16[10] = level
    hf[6], hf[7], hf[8] = scene_center
    hf[9], hf[10] = lo, hi
    sh_ = np.zeros(SECTION_HEADER_BYTES, np.uint8)
    s32, s16, sf = sh_.view(np.uint32), sh_.view(np.uint16), sh_.view(np.float32)
    s32[0], s32[1] = n, n
    if level >= 1:
        s32[2], s32[3] = bucket_size, len(bucket_centres)
        sf[4] = block_size
        s16[10] = BUCKET_STORAGE_BYTES
        s32[6] = L["scale_range"]
        s32[8], s32[9] = full_count, len(partial_lengths)
    s32[7] = len(section)
    s16[20] = sh_degree
    return header.tobytes() + sh_.tobytes() + section

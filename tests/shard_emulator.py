"""CPU emulation of what one rank exports in the multi-GPU shard protocol (fq_shard.cu / fq_layout.h):
written directly from the block semantics, independent of the CUDA code, so that the host-side combine
step (fqgpu_shard_combine_host) and the collective plumbing can be tested without a GPU."""
import numpy as np

POS_BINS = 512
OFF_HIST_SEQ, OFF_HIST_QUAL = 0, 256
OFF_SEQ_LEN = 512
OFF_QUAL_LEN = OFF_SEQ_LEN + POS_BINS + 1
OFF_SEQ_LOG2 = OFF_QUAL_LEN + POS_BINS + 1
OFF_POS_SUM = OFF_SEQ_LOG2 + 64
OFF_SEQ_LEN_MIN = OFF_POS_SUM + POS_BINS + 1
OFF_SEQ_LEN_MAX, OFF_QUAL_LEN_MIN, OFF_QUAL_LEN_MAX = OFF_SEQ_LEN_MIN + 1, OFF_SEQ_LEN_MIN + 2, OFF_SEQ_LEN_MIN + 3
BLOCK_WORDS = ((OFF_QUAL_LEN_MAX + 1 + 31) // 32) * 32
SH_OFF_HEAD_POS = 0
SH_OFF_SCALARS = POS_BINS + 1
(SH_LINES, SH_BYTES, SH_OPEN_LEN, SH_LAST_BYTE, SH_FIRST_BYTE, SH_HEAD_LEN, SH_HEAD_CR, SH_HYP, SH_HYP_VALID, SH_EXACT,
 SH_META_LINES, SH_META_QMIN, SH_META_QMAX, SH_META_STATUS, SH_META_PENDING_CR, SH_META_CUR_HAS, SH_META_CUR_MIN,
 SH_META_CUR_MAX, SH_PRESENT, SH_NSCALARS) = range(20)
SHARD_EXTRA_WORDS = ((SH_OFF_SCALARS + SH_NSCALARS + 31) // 32) * 32
SHARD_WORDS = BLOCK_WORDS + SHARD_EXTRA_WORDS
U64_MAX = (1 << 64) - 1


def make_block(shard: bytes, rank: int, hyp: int, meta_state=None) -> np.ndarray:
    """Block of a hypothesis shard (rank > 0) or of rank 0 (hyp = 0, exact)."""
    blk = [0] * SHARD_WORDS
    blk[OFF_SEQ_LEN_MIN] = blk[OFF_QUAL_LEN_MIN] = U64_MAX
    ex = BLOCK_WORDS
    sc = ex + SH_OFF_SCALARS
    n = len(shard)
    # content bytes: everything except '\n', except a '\r' directly before a '\n', except a final '\r'
    # (its fate depends on the next shard / the end of the stream)
    line = 0      # shard-relative line index
    pos = 0       # position inside the current line, relative to the shard start for line 0
    first_nl = shard.find(b"\n")
    for o, b in enumerate(shard):
        cls = (hyp + line) & 3
        if b == 0x0A:
            if line > 0 or rank == 0:
                cr = 1 if (pos > 0 and shard[o - 1] == 0x0D) else 0
                L = pos - cr
                if cls == 1:
                    blk[OFF_SEQ_LEN + min(L, POS_BINS)] += 1
                    blk[OFF_SEQ_LOG2 + L.bit_length()] += 1
                    blk[OFF_SEQ_LEN_MIN] = min(blk[OFF_SEQ_LEN_MIN], L); blk[OFF_SEQ_LEN_MAX] = max(blk[OFF_SEQ_LEN_MAX], L)
                elif cls == 3:
                    blk[OFF_QUAL_LEN + min(L, POS_BINS)] += 1
                    blk[OFF_QUAL_LEN_MIN] = min(blk[OFF_QUAL_LEN_MIN], L); blk[OFF_QUAL_LEN_MAX] = max(blk[OFF_QUAL_LEN_MAX], L)
            line += 1
            pos = 0
            continue
        content = True
        if b == 0x0D:
            content = o + 1 < n and shard[o + 1] != 0x0A
        if content and (cls & 1):
            blk[(OFF_HIST_QUAL if cls == 3 else OFF_HIST_SEQ) + b] += 1
            if cls == 3:
                if line == 0 and rank > 0:
                    blk[ex + SH_OFF_HEAD_POS + min(pos, POS_BINS)] += b   # detached head: relative positions
                else:
                    blk[OFF_POS_SUM + min(pos, POS_BINS)] += b
        pos += 1
    blk[sc + SH_LINES] = line
    blk[sc + SH_BYTES] = n
    blk[sc + SH_OPEN_LEN] = pos if line else n
    blk[sc + SH_LAST_BYTE] = shard[-1] if n else 0
    blk[sc + SH_FIRST_BYTE] = shard[0] if n and rank > 0 else 0
    if rank > 0 and first_nl >= 0:
        blk[sc + SH_HEAD_LEN] = first_nl
        blk[sc + SH_HEAD_CR] = 1 if (first_nl > 0 and shard[first_nl - 1] == 0x0D) else 0
    blk[sc + SH_HYP] = hyp if rank > 0 else 0
    blk[sc + SH_HYP_VALID] = 1 if rank > 0 else 0
    blk[sc + SH_EXACT] = 1 if rank == 0 else 0
    blk[sc + SH_META_QMIN] = blk[sc + SH_META_QMAX] = U64_MAX  # -1
    blk[sc + SH_PRESENT] = 1
    return np.array(blk, dtype=np.uint64)

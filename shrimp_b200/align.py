"""Alignment strings from the edit script of a shrimp_hit: pretty_print of common/sw-full-ls.c:524-560
(dbalign / qralign) and the SAM fields gmapper/output.c derives from them (make_cigar :15-65,
hit_output :470-700).  Host-side formatting only -- the unchanged output.c does this in a real
integration; here it lets tests compare against the reference's SAM."""
from __future__ import annotations

import numpy as np

_LS_LETTERS = b"ACGTUMRWSYKVHDBN"      # base_translate, fasta.c:694-696


def align_strings(edit: np.ndarray, genome_codes: np.ndarray, genome_start: int, read_codes: np.ndarray,
                  read_start: int):
    """genome_codes: codes of the strand the hit is on (contig coordinates), read_codes: the read as given."""
    db, qr = bytearray(), bytearray()
    gi, ri = genome_start, read_start
    for op in edit:
        op = int(op) & 3
        if op == 2:      # BACK_DELETION: read base against a gap
            db += b"-"
            qr.append(_LS_LETTERS[int(read_codes[ri]) & 15])
            ri += 1
        elif op == 1:    # BACK_INSERTION: genome base against a gap
            db.append(_LS_LETTERS[int(genome_codes[gi]) & 15])
            qr += b"-"
            gi += 1
        else:
            db.append(_LS_LETTERS[int(genome_codes[gi]) & 15])
            qr.append(_LS_LETTERS[int(read_codes[ri]) & 15])
            gi += 1
            ri += 1
    return bytes(db), bytes(qr)


def cigar_from_edit(edit: np.ndarray, read_start0: int, rmapped: int, read_len: int, reverse: bool,
                    clip: str = "S") -> str:
    """make_cigar + reverse_cigar (output.c:15-90): BACK_INSERTION columns are SAM 'D', BACK_DELETION 'I'."""
    ops = []
    if read_start0 + 1 > 1:
        ops.append((read_start0, clip))
    prev, run = None, 0
    for op in edit:
        c = {1: "D", 2: "I", 3: "M"}[int(op) & 3]
        if c == prev:
            run += 1
        else:
            if prev is not None:
                ops.append((run, prev))
            prev, run = c, 1
    if prev is not None:
        ops.append((run, prev))
    read_end = read_start0 + rmapped
    if read_end != read_len:
        ops.append((read_len - read_end, clip))
    if reverse:
        ops = ops[::-1]
    return "".join(f"{n}{o}" for n, o in ops)


def sam_fields(hit, edit: np.ndarray, read_len: int, genome_len_cn: int, colour_space: bool = False):
    """(flag, cn, pos, cigar, AS, NM) as hit_output prints them for an unpaired read (output.c:470-700)."""
    reverse = int(hit["gen_st"]) == 1
    read_start = int(hit["read_start"]) + 1
    read_end = read_start + int(hit["rmapped"]) - 1
    if not reverse:
        pos = int(hit["genome_start"]) + 1
    else:
        right = genome_len_cn - int(hit["genome_start"])
        pos = right - (read_end - read_start - int(hit["deletions"]) + int(hit["insertions"]))
    cigar = cigar_from_edit(edit, int(hit["read_start"]), int(hit["rmapped"]), read_len, reverse,
                            "H" if colour_space else "S")
    nm = int(hit["mismatches"]) + int(hit["deletions"]) + int(hit["insertions"])
    return (16 if reverse else 0, int(hit["cn"]), pos, cigar, int(hit["score_full"]), nm, int(hit["gmapped"]))


def cs_layer_letters(colours: np.ndarray, initbp: int) -> np.ndarray:
    """[4, n] letter translations of a colour read, layer k starting from letter (k + initbp) % 4; a colour N
    gives letter N and restarts the layer (sw-full-cs.c:1181-1196, cstols util.h:157-180)."""
    n = len(colours)
    out = np.zeros((4, n), dtype=np.uint8)
    for k in range(4):
        letter = (k + initbp) % 4
        for j in range(n):
            c = int(colours[j])
            if c == 15:
                out[k, j] = 15
                letter = (k + initbp) % 4
            else:
                r = 15 if (letter == 15 or c > 3) else ((4 + letter + c) % 4 if letter % 2 == 0 else (4 + letter - c) % 4)
                out[k, j] = r
                letter = r
    return out


def align_strings_cs(edit: np.ndarray, genome_codes: np.ndarray, genome_start: int, colours: np.ndarray, initbp: int,
                     read_start: int):
    """pretty_print of common/sw-full-cs.c:945-1060 from the edit script: the letter shown is the one of the layer
    in bits 4-5, lower case on a crossover column; an aligned N takes the genome's letter."""
    qr = cs_layer_letters(colours, initbp)
    db, q = bytearray(), bytearray()
    gi, ri = genome_start, read_start
    for op in edit:
        op = int(op)
        ty, xo, k = op & 3, op & 4, (op >> 4) & 3
        if ty == 2:
            c = _LS_LETTERS[int(qr[k, ri]) & 15]
            ri += 1
            db += b"-"
            q.append(c | 0x20 if xo else c)
        elif ty == 1:
            db.append(_LS_LETTERS[int(genome_codes[gi]) & 15])
            q += b"-"
            gi += 1
        else:
            c = _LS_LETTERS[int(qr[k, ri]) & 15]
            g = _LS_LETTERS[int(genome_codes[gi]) & 15]
            ri += 1
            gi += 1
            db.append(g)
            c = c | 0x20 if xo else c
            if c in (ord("n"), ord("N")):
                c = g | 0x20 if xo else g
            q.append(c)
    return bytes(db), bytes(q)


def post_sw_seq_qual(edit: np.ndarray, quals: np.ndarray, reverse: bool, have_quals: bool):
    """SEQ / QUAL columns of a colour-space alignment with mapping qualities (gmapper/output.c:486-546, :607-628)
    from what shrimp_gpu_map_reads returns after post_sw: edit bytes with bit 3 set carry the corrected base call in
    bits 4-5; `quals` are the rmapped bytes of sfrp->qual that follow the edit script."""
    rc = bytes.maketrans(b"ACGT", b"TGCA")
    seq = bytes(_LS_LETTERS[(int(op) >> 4) & 3] for op in edit if (int(op) & 3) != 1)
    q = bytes(quals.tobytes()) if have_quals else b"*"
    if reverse:
        seq = seq.translate(rc)[::-1]
        if have_quals:
            q = q[::-1]
    return seq.decode(), q.decode()

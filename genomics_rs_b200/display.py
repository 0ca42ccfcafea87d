"""`impl Display for AlignedSequences` -- host-side restatement of /root/reference/src/alignment/display.rs:9-127."""
from __future__ import annotations

import logging
from decimal import Decimal

log = logging.getLogger("genomics_rs_b200")

DISP_MAX_WIDTH = 200  # display.rs:7

_INSERTS = (2, 4)   # Insert, OpenInsert
_DELETES = (3, 5)   # Delete, OpenDelete


def _rust_f64(x: float) -> str:
    """Rust `{}` for f64: shortest round-trip digits, never an exponent, no trailing '.0'; NaN -> 'NaN'."""
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "inf" if x > 0 else "-inf"
    s = format(Decimal(repr(x)), "f")
    if "." in s:
        s = s.rstrip("0").rstrip(".")
    return s


def _pct2(num: int, den: int) -> str:
    """Rust `{:.2}` of (num as f64 / den as f64) * 100.0"""
    if den == 0:
        return "NaN"
    return f"{(num / den) * 100.0:.2f}"


def format_alignment(a) -> str:
    s1, s2 = a.s1.sequence.encode("utf-8"), a.s2.sequence.encode("utf-8")
    if len(s1) <= DISP_MAX_WIDTH and len(s2) <= DISP_MAX_WIDTH:      # display.rs:12-18 (log side effects only)
        log.info("Original Sequences:")
        log.info("%s", a.s1)
        log.info("%s", a.s2)
    else:
        log.warning("Sequences are too long to display.")
    out = []
    s1_out, align_out, s2_out = [], [], []
    s1_idx = s2_idx = 0
    horizontal_len = 0
    align_idx = 0
    ops = a.ops[::-1]                                                # display.rs:26: iter().rev()
    n = len(ops)
    sym = {0: "|", 1: "x", 2: " ", 3: " ", 4: "%", 5: "%"}
    while align_idx < n:                                             # display.rs:31
        c = int(ops[align_idx])
        if horizontal_len > DISP_MAX_WIDTH:                          # display.rs:34-44
            out.append(f"\n\n{align_idx - DISP_MAX_WIDTH}-{align_idx}:\n\n")
            out.append("".join(s1_out) + "\n" + "".join(align_out) + "\n" + "".join(s2_out) + "\n")
            s1_out, align_out, s2_out = [], [], []
            horizontal_len = 0
        if c in _INSERTS:                                            # display.rs:47-57
            s1_out.append("-")
        elif s1_idx < len(s1):
            s1_out.append(chr(s1[s1_idx]))
            s1_idx += 1
        align_out.append(sym[c])                                     # display.rs:60-69
        if c in _DELETES:                                            # display.rs:72-82
            s2_out.append("-")
        elif s2_idx < len(s2):
            s2_out.append(chr(s2[s2_idx]))
            s2_idx += 1
        horizontal_len += 1
        align_idx += 1
    s1_str = "".join(s1_out)
    out.append(f"\n\n{align_idx - len(s1_str.encode('utf-8'))}-{align_idx}:\n\n")   # display.rs:88
    out.append(s1_str + "\n" + "".join(align_out) + "\n" + "".join(s2_out) + "\n")
    out.append(f"\n\nAlignment Score: {a.score}\n")                  # display.rs:92
    out.append(f"Matches: {a.matches}/{align_idx} ({_pct2(a.matches, align_idx)}%)\n")
    out.append(f"Mismatches: {a.mismatches}/{align_idx} ({_pct2(a.mismatches, align_idx)}%)\n")
    out.append(f"Gap Extensions: {a.gap_extensions}/{align_idx} ({_pct2(a.gap_extensions, align_idx)}%)\n")
    out.append(f"Opening Gaps: {a.opening_gaps}/{align_idx} ({_pct2(a.opening_gaps, align_idx)}%)\n")
    ident = (a.matches / align_idx) * 100.0 if align_idx else float("nan")
    out.append(f"Percent Identity {_rust_f64(ident)}%\n")            # display.rs:120-125
    return "".join(out)


# ---------------------------------------------------------------------------------------------------------------
# print_alignment_table / print_scores_table -- display.rs:131-220 (called by retrace, algo.rs:438)
_NEG_INF_PRINT = -9223372036854775700     # display.rs:213


def _ansi(txt: str, code: str, bold: bool, color: bool) -> str:
    if not color:
        return txt
    return f"\x1b[{'1;' if bold else ''}{code}m{txt}\x1b[0m"


def format_scores_table(plane) -> str:
    """display.rs:190-220 for one plane ((m+1) x (n+1) int64)"""
    rows, cols = plane.shape
    out = [". \t" + "".join(f"{j}\t" for j in range(cols)) + "\n"]
    for i in range(rows):
        vals = "".join(("-inf" if int(v) <= _NEG_INF_PRINT else str(int(v))) + "\t" for v in plane[i])
        out.append(f"{i}\t{vals}\n")
    return "".join(out)


def format_alignment_table(a, planes, color: bool = False):
    """The text print_alignment_table writes to stdout (display.rs:131-188), or None when the reference skips it
    (s1 >= 200 or s2 >= 2000 characters: "Sequence table too large to visualize").  `planes` = (insert, delete, sub)
    from genomics_rs_b200.score_planes().  `color` adds the ANSI colours the reference's `colored` crate emits on a
    terminal."""
    s1, s2 = a.s1.sequence, a.s2.sequence
    if not (len(s1.encode("utf-8")) < DISP_MAX_WIDTH and len(s2.encode("utf-8")) < DISP_MAX_WIDTH * 10):   # display.rs:139
        log.warning("Sequence table too large to visualize")
        return None
    log.info("Computing sequence table visualization...")
    path = {}
    for c, i, j in a.alignment:                  # .find(): the FIRST entry of the walk at (i+1, j+1) wins
        path.setdefault((i, j), int(c))
    glyph = {0: ("M", "32", False), 1: ("X", "31", False), 2: ("I", "34", False), 3: ("D", "36", False),
             4: ("I", "34", True), 5: ("D", "36", True)}
    out = ["\nSequence Table (S1 columns, S2 rows):\n\n", " " + s2 + "\n"]
    for i, ch in enumerate(s1):
        row = [ch]
        for j in range(len(s2)):
            c = path.get((i + 1, j + 1))
            row.append("." if c is None else _ansi(*glyph[c][:2], glyph[c][2], color))
        out.append("".join(row) + "\n")
    ins, dele, sub = planes
    out.append("Delete Scores\n" + format_scores_table(dele))
    out.append("Insert Scores\n" + format_scores_table(ins))
    out.append("Sub Scores\n" + format_scores_table(sub))
    return "".join(out)

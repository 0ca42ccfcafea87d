"""Sequence / SequenceContainer -- host-side mirror of /root/reference/src/sequence.rs."""
from __future__ import annotations

import logging
from dataclasses import dataclass, field
from typing import List, Optional

log = logging.getLogger("genomics_rs_b200")


@dataclass
class Sequence:
    """sequence.rs:9-12; Display at sequence.rs:14-18."""
    name: str
    sequence: str

    def __str__(self) -> str:
        return f"{self.name}: {self.sequence}"

    def bytes(self) -> bytes:
        return self.sequence.encode("utf-8")


def _opt_byte(b: bytes, i: int) -> Optional[int]:
    return b[i] if 0 <= i < len(b) else None


@dataclass
class SequenceContainer:
    """sequence.rs:32-34"""
    sequences: List[Sequence] = field(default_factory=list)

    def from_fasta(self, filepath: str) -> None:
        """sequence.rs:45-95: '>' opens a record (name = rest, trimmed); other non-empty lines are trimmed
        and appended to the last record; data before the first header is warned about and dropped;
        an unreadable file logs an error and adds nothing."""
        sequences: List[Sequence] = []
        have_header = False
        try:
            with open(filepath, "rb") as fh:
                raw = fh.read()
        except OSError:
            log.error("Could not open file: %s", filepath)
            raw = None
        if raw is not None:
            for bline in raw.split(b"\n"):
                try:
                    line = bline.decode("utf-8")      # io::Lines yields Err on invalid UTF-8; map_while stops there
                except UnicodeDecodeError:
                    break
                if line.endswith("\r"):
                    line = line[:-1]                  # BufRead::lines strips "\n" and "\r\n"
                if not line:
                    continue
                if line.startswith(">"):
                    name = line[1:].strip()
                    log.info("Sequence Found (ID: %d): %s", len(self.sequences) + len(sequences), filepath)
                    sequences.append(Sequence(name=name, sequence=""))
                    have_header = True
                elif have_header:
                    sequences[-1].sequence += line.strip()
                else:
                    log.warning("Sequence data found without a header")
        log.debug("Loaded %d sequences", len(sequences))
        self.sequences.extend(sequences)

    def from_fasta_dir(self, fasta_dir: str) -> None:
        """Directory ingestion of the reference's `compare` sub-command (main.rs:227-239): every file whose extension is
        `fasta`, all of its records.  Files are taken in SORTED name order: read_dir order is unspecified and NW with the
        reference's tie-breaks is not symmetric, so the order is part of the contract."""
        import os
        for name in sorted(os.listdir(fasta_dir)):
            if name.rsplit(".", 1)[-1] == "fasta" and "." in name:
                self.from_fasta(os.path.join(fasta_dir, name))

    def is_match(self, i: int, j: int, reverse_sequences: bool = False) -> bool:
        """sequence.rs:102-115.  Option<u8> equality: both out of range compares equal (None == None)."""
        s1 = self.sequences[0].bytes()
        s2 = self.sequences[1].bytes()
        if reverse_sequences:               # sequence.rs:103-112 (lengths are swapped there; dead code in the reference)
            i = len(s2) - i
            j = len(s1) - j
        return _opt_byte(s1, i) == _opt_byte(s2, j)

"""Scoring configuration -- host-side mirror of /root/reference/src/config.rs:1-40 and config.toml:1-5."""
from __future__ import annotations

import logging
import sys
import tomllib
from dataclasses import dataclass

log = logging.getLogger("genomics_rs_b200")


@dataclass(frozen=True)
class Scores:
    """config.rs:6-13.  All four are i64 in the reference; the GPU path range-checks them into int32."""
    s_match: int
    s_mismatch: int
    g: int
    h: int

    def as_tuple(self):
        return (self.s_match, self.s_mismatch, self.g, self.h)


@dataclass(frozen=True)
class Config:
    """config.rs:15-18"""
    scores: Scores


def parse_config(text: str) -> Config:
    data = tomllib.loads(text)
    sc = data["scores"]
    vals = {}
    for key in ("s_match", "s_mismatch", "g", "h"):
        v = sc[key]
        if isinstance(v, bool) or not isinstance(v, int):
            raise ValueError(f"scores.{key} must be an integer")
        vals[key] = v
    return Config(scores=Scores(**vals))


def get_config(filepath: str) -> Config:
    """config.rs:21-40: read + parse; on either failure log an error and exit(1) (config.rs:26,34)."""
    try:
        with open(filepath, "r", encoding="utf-8") as fh:
            contents = fh.read()
    except (OSError, UnicodeDecodeError):
        log.error("Could not read config file: %s", filepath)
        sys.exit(1)
    try:
        return parse_config(contents)
    except Exception:  # toml syntax, missing table/key, wrong type
        log.error("Could not parse config file: %s", filepath)
        sys.exit(1)

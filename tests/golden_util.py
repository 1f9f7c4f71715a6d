"""Helpers that rebuild the inputs of the committed golden vectors (tests/golden/*.json)."""
import json
import os

import numpy as np

import orc

HERE = os.path.dirname(os.path.abspath(__file__))
CODE = {"u8": orc.U8, "u16": orc.U16, "u32": orc.U32, "u64": orc.U64,
        "i8": orc.I8, "i16": orc.I16, "i32": orc.I32, "i64": orc.I64}


def load(name):
    with open(os.path.join(HERE, "golden", name + ".json")) as f:
        return json.load(f)


def small_input(c):
    dt = orc.NP_OF[CODE[c["dtype"]]]
    if "iota" in c:
        lo, n = c["iota"]
        return np.arange(lo, lo + n, dtype=dt)
    return np.array(c["values"], dtype=dt)


def large_input(c):
    dt = CODE[c["dtype"]]
    if c["gen"] == "kat_fill":
        a = orc.kat_fill(dt, c["n"], c["seed"])
        return (a >> 33).astype(np.int64) if dt == orc.I64 else a   # see make_golden.py (C11)
    return orc.synth_frame(dt, c["width"], c["height"], c["lambda"], c["peaks"], c["seed"],
                           20.0, c.get("amp_hi", 3000.0))

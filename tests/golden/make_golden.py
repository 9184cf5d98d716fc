"""Extracts the reference's golden attention tables into attention_tables.json.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
Sources (literal tables inside the reference's own tests):
  * csrc/tests/flash_attn_tests.rs:54-70   non-causal, shape (3 heads, 2 tokens, 8 dims), fp16
    (the same table is asserted at :75-88, :119-135 and :217-233 through three different APIs)
  * models/src/flash_attention.rs:688-704  causal, same inputs
The fixture also records how the reference builds its inputs: `Tensor::arange(0u32, 48)` cast to
f16, reshaped, and divided with candle's affine op, i.e. multiplied by f16(1/c).
"""
import json
import re
from pathlib import Path

REF = Path("/root/reference")


def tables_in(path: Path):
    src = path.read_text()
    out = []
    # a table = "&[" followed by 3 x 2 rows of 8 floats
    for m in re.finditer(r"&\[\s*((?:\[\s*(?:\[[^\[\]]*\],?\s*)+\],?\s*)+)\]", src):
        rows = re.findall(r"\[([^\[\]]+)\]", m.group(1))
        vals = [[float(x) for x in r.replace("\n", " ").split(",") if x.strip()] for r in rows]
        if len(vals) == 6 and all(len(r) == 8 for r in vals):
            line = src[: m.start()].count("\n") + 1
            out.append((line, [vals[0:2], vals[2:4], vals[4:6]]))
    return out


def main():
    acausal = tables_in(REF / "csrc/tests/flash_attn_tests.rs")
    causal = tables_in(REF / "models/src/flash_attention.rs")
    assert len(acausal) == 4 and all(t[1] == acausal[0][1] for t in acausal), "expected 4 identical tables"
    assert len(causal) == 1
    fixture = {
        "inputs": {
            "recipe": "a = arange(48) as f16, reshaped (1,3,2,8) = (batch, heads, tokens, dim); "
                      "q = a * f16(1/30), k = a * f16(1/40), v = a * f16(1/50); softmax_scale = 0.5",
            "numel": 48, "heads": 3, "tokens": 2, "dim": 8, "dtype": "f16",
            "q_div": 30.0, "k_div": 40.0, "v_div": 50.0, "softmax_scale": 0.5,
        },
        "acausal": {"source": "csrc/tests/flash_attn_tests.rs", "lines": [t[0] for t in acausal],
                    "layout": "[head][token][dim], rounded to 4 decimals", "table": acausal[0][1]},
        "causal": {"source": "models/src/flash_attention.rs", "lines": [causal[0][0]],
                   "layout": "[head][token][dim], rounded to 4 decimals", "table": causal[0][1]},
    }
    out = Path(__file__).with_name("attention_tables.json")
    out.write_text(json.dumps(fixture, indent=1) + "\n")
    print("wrote", out)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Builds tests/golden/c1_sample.json from the reference's shipped sample
(/root/reference/portfolios.json, /root/reference/macro.json): the ten instrument/weight
lists and the 10 x 10 macro-factor rows, i.e. BASELINE config C1's inputs, so that the
GPU box (which has no /root/reference) can run the shipped-sample parity test.
Run here only:  python tests/golden/make_c1_fixture.py
"""
import json
import os
import sys

REF = "/root/reference"


def main():
    if not os.path.isdir(REF):
        sys.exit("needs /root/reference")
    ports = [json.loads(l) for l in open(os.path.join(REF, "portfolios.json")) if l.strip()]
    macro = [json.loads(l) for l in open(os.path.join(REF, "macro.json")) if l.strip()]
    out = {
        "source": "portfolios.json:1-10 and macro.json:1-10 of mapr-demos/monte-carlo-portfolios",
        "portfolios": [{"id": p.get("id"), "instruments": [x["instrument"] for x in p["instruments"]],
                        "weights": [x["weight"] for x in p["instruments"]]} for p in ports],
        "macro": [[row[f"m{f}"] for f in range(10)] for row in sorted(macro, key=lambda r: r["date_offset"])],
        # setup.txt:47-48: the stored document of portfolio 0
        "kat": {"id": "0", "uncompressedIntegerSize": 17, "algo": "bpacking+variablebyte",
                "compressed_b64": "AAAAABaFhU8AoBCfiQvOmoIHgQeORMPDeYo+y4lApogAAAAA",
                "min_instrument": 5, "max_instrument": 4112,
                "min_weight": 11.108738657302295, "max_weight": 299.2617711307142},
    }
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_sample.json")
    json.dump(out, open(dst, "w"), separators=(",", ":"))
    print("wrote", dst, [len(p["instruments"]) for p in out["portfolios"]])


if __name__ == "__main__":
    main()

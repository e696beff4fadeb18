#!/usr/bin/env python
"""Generates tests/golden/aux_parser.json: what the REFERENCE's own read_inputfile<T>()
(src/acquisition.cxx:31-59, compiled where it lies into oracle/_ref/ref_read_inputfile by
oracle/Makefile) returns for a set of window/baseline text files, including its corner cases
(a failed extraction zeroes the kept value: "0.5 abc" -> 0).  Run in the build container:
    python tests/golden/make_aux_golden.py
"""
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
EXE = os.path.join(ROOT, "oracle", "_ref", "ref_read_inputfile")

CASES = {
    "plain": "0.0\n0.25\n0.5\n0.75\n1.0\n0.75\n0.5\n0.25\n",
    "columns_last_wins": "100 0.5\n200 300 0.25\n1 2 3\n",
    "comments_and_blanks": "# header\n\n   # indented\n1.5\n\t\n2.5 # trailing text\n#3.5\n",
    "trailing_garbage_zeroes": "0.5 abc\n1e3 4x\n7 \nxyz\n-2.5e-3\n.5\n5.\n",
    "signs_exponents": "+1\n-1\n1e-40\n1e40\n0x10\n1,5\n 3 ;\n",
    "crlf": "1.0\r\n2.0\r\n",
    "no_final_newline": "4\n5",
    "special": "inf\nnan\n-inf\n1e400\n",
}


def run(kind, text):
    r = subprocess.run([EXE, kind], input=text, capture_output=True, text=True, check=True)
    return r.stdout.split()


def main():
    out = {}
    for name, text in CASES.items():
        out[name] = {"text": text, "float": run("float", text), "double": run("double", text)}
    with open(os.path.join(HERE, "aux_parser.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    for k, v in out.items():
        print(k, v["double"])


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Regenerates tests/golden/ from the reference tree (run in the build container only).

Copies the reference's own known-answer DATA (not source):
  * conformance/valid/*.zxc + .expected + .zxd   -> tests/golden/valid/     (decode KAT)
  * conformance/invalid/*.zxc                    -> tests/golden/invalid/   (reject KAT)
    with the pinned error table of conformance/test_conformance.c:228-249
    transcribed to tests/golden/invalid/expected.json
  * tests/format/golden/*.zxc + golden.sha256    -> tests/golden/format/    (encoder KAT)
and generates seeded differential fixtures with the UNMODIFIED reference library
(oracle/_ref/libzxc_ref.so): tests/golden/diff/*.bin (input) + *.zxc (frame).
/root/reference does not exist on the GPU box; tests read only tests/golden/.
"""
import json, os, shutil, sys, glob, ctypes
import numpy as np

REF = os.environ.get("ZXC_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

INVALID = {  # conformance/test_conformance.c:228-249
    "all_0xff_garbage": -4, "bad_block_checksum": -7, "bad_block_size_field": -14,
    "bad_block_type": -13, "bad_checksum_algo": -6, "bad_enc_lit": -8, "bad_eof_compsize": -6,
    "bad_header_crc": -6, "bad_magic": -4, "bad_version": -5, "corrupt_payload": -7,
    "dict_required": -15, "ghi_forged_offset": -9, "glo_forged_enc_off": -8,
    "glo_insufficient_slack": -8, "magic_then_zeros": -5, "too_short_4bytes": -3,
    "truncated_header_only": -3, "truncated_mid_block": -3, "zero_length": -3,
}


def copy_tree(src_glob, dst):
    os.makedirs(dst, exist_ok=True)
    for f in sorted(glob.glob(src_glob)):
        shutil.copy(f, dst)


def main():
    copy_tree(f"{REF}/conformance/valid/*", f"{HERE}/valid")
    copy_tree(f"{REF}/conformance/invalid/*.zxc", f"{HERE}/invalid")
    with open(f"{HERE}/invalid/expected.json", "w") as f:
        json.dump(INVALID, f, indent=1, sort_keys=True)
    copy_tree(f"{REF}/tests/format/golden/*.zxc", f"{HERE}/format")
    with open(f"{REF}/tests/format/golden.sha256") as f, open(f"{HERE}/format/golden.sha256", "w") as g:
        for line in f:
            h, p = line.split()
            g.write(f"{h}  {os.path.basename(p)}\n")
    print("fixtures copied")


if __name__ == "__main__":
    main()

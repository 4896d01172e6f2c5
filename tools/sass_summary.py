"""profiles/sass_summary.txt: per translation unit of librs_b200.so, how often the SASS mnemonics that prove the Blackwell
paths occur (cuobjdump -sass on the objects the current build linked; see recommendsystem_b200/build.py for the stamp).
usage: python tools/sass_summary.py > profiles/sass_summary.txt"""
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLS = [("UTC.MMA", r"UTC[A-Z]*MMA"), ("UTMALDG", r"UTMALDG"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
        ("UTCBAR", r"UTCBAR"), ("LDG.256", r"LDG\.[A-Z0-9.]*256"), ("HMMA", r"(^|[^A-Z])HMMA")]


def main():
    newest = {}
    for f in glob.glob(os.path.join(ROOT, "build", "obj", "*.o")):
        unit = os.path.basename(f).rsplit(".", 2)[0]
        if unit not in newest or os.path.getmtime(f) > os.path.getmtime(newest[unit]):
            newest[unit] = f
    print("# SASS evidence (cuobjdump -sass build/obj/<unit>.o | grep -c <mnemonic>), round 2 final objects, sm_100a")
    print("# UTC*MMA = tcgen05.mma, UTMALDG = TMA loads, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, "
          "LDG.256 = 32-byte global loads,")
    print("# HMMA = legacy mma.sync / wmma (counted as a whole mnemonic, UTCHMMA excluded): must be 0")
    print(f"{'unit':28s}" + "".join(f"{c:>9s}" for c, _ in COLS))
    for unit in sorted(newest):
        sass = subprocess.run(["cuobjdump", "-sass", newest[unit]], capture_output=True, text=True).stdout
        lines = sass.splitlines()
        print(f"{unit:28s}" + "".join(f"{sum(1 for ln in lines if re.search(rx, ln)):9d}" for _, rx in COLS))


if __name__ == "__main__":
    main()

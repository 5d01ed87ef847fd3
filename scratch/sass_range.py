"""Print the SASS of one kernel between two addresses: python scratch/sass_range.py <lib.so> <mangled-name> <from-hex> <to-hex>"""
import re, subprocess, sys
lib, name, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
out = subprocess.run(["cuobjdump", "-sass", "-fun", name, lib], capture_output=True, text=True).stdout
for line in out.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m and lo <= int(m.group(1), 16) <= hi:
        print("%05x  %s" % (int(m.group(1), 16), m.group(2)))

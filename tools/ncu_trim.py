"""Keep only the executed instructions of an ncu SASS source page (CSV on stdin -> stdout); the full page of the NTT kernels is
tens of megabytes because every butterfly policy is in the binary and one of them runs."""
import csv
import sys
r = csv.reader(sys.stdin)
w = csv.writer(sys.stdout)
rows = list(r)
w.writerow(rows[0])
hdr = rows[1]
w.writerow(hdr)
ix = hdr.index("Instructions Executed")
for row in rows[2:]:
    if len(row) == len(hdr) and row[ix] not in ("0", ""):
        w.writerow(row)

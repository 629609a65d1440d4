#!/bin/bash
# Per-kernel histogram of the tensor-core / tensor-memory / TMA opcodes in the built library (run where it was built; no GPU needed).
SO=${1:-video-summarization_b200/vsum_b200/libvsum_b200.so}
cuobjdump -sass "$SO" | awk '
  /Function :/ { fn=$3 }
  { while (match($0, /UTC[A-Z0-9]*MMA[.A-Z0-9_]*|LDTM[.A-Za-z0-9_]*|STTM[.A-Za-z0-9_]*|UTMALDG[.A-Z0-9_]*|UTMASTG[.A-Z0-9_]*|UTMAPF[.A-Z0-9_]*|UTCBAR[.A-Z0-9_]*|UTCATOMSWS[.A-Z0-9_]*|MUFU\.EX2[.A-Z0-9_]*|SYNCS[.A-Z0-9_]*/)) {
      op=substr($0, RSTART, RLENGTH); c[fn SUBSEP op]++; $0=substr($0, RSTART+RLENGTH) } }
  END { for (k in c) { split(k, a, SUBSEP); print a[1] "\t" a[2] "\t" c[k] } }' | sort | c++filt | \
  awk -F'\t' '{ n=$1; gsub(/\(anonymous namespace\)::/, "", n); sub(/\(.*/, "", n); if (n != last) { print ""; print n; last=n } printf "    %-28s %d\n", $2, $3 }'

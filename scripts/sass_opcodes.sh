#!/bin/bash
# Opcode histogram of the sm_100a cubins inside vod_b200/libvodb.so (cuobjdump -sass): the Blackwell-native evidence
# (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA loads, UTCBAR = tcgen05.commit) next to what is absent
# (HMMA = mma.sync, HGMMA = wgmma). Usage: scripts/sass_opcodes.sh > profiles/<round>_sass_opcodes.txt
set -eu
cd "$(dirname "$0")/.."
SO=vod_b200/libvodb.so
echo "# $(date -u +%FT%TZ) $(sha256sum $SO | cut -c1-16) $SO"
cuobjdump -lelf $SO | sed 's/^/# /'
echo "# --- per kernel: tcgen05 / TMA / TMEM opcodes"
cuobjdump -sass $SO | awk '
  /Function :/ { fn=$3; sub(/^_ZN4vodb[0-9]+_GLOBAL__N__[0-9a-f]+_[0-9]+_[a-z_]+_cu_[0-9a-f]+/,"",fn) }
  /^[ \t]+\/\*[0-9a-f]+\*\// { op=$2; if (op ~ /^@/) op=$3; sub(/;$/,"",op); split(op,a,"."); base=a[1];
    if (op ~ /^UTC|^LDTM|^STTM|^UTMA|^UBLKCP|^HMMA|^HGMMA|^UTCBAR|^UTCATOM/) k[fn" "op]++ ; all[base]++ }
  END { for (x in k) print k[x], x | "sort -k2,2 -k1,1nr"; close("sort -k2,2 -k1,1nr");
        print "# --- whole library: opcode histogram (base mnemonic)"; for (x in all) print all[x], x | "sort -k1,1nr" }'

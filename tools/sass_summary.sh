#!/bin/bash
# Per-kernel counts of the Blackwell-specific SASS mnemonics in the built library (cuobjdump runs without a GPU):
# UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (TMEM), UTMALDG/UTMASTG = TMA tensor loads/stores, UTCBAR = tcgen05.commit,
# .2CTA = cta_group::2 forms (MMA / TMA / commit), UBLKCP = 1-D bulk copies (cp.async.bulk), LDG.256 = 256-bit global loads, F*2 = packed fp32.
set -e
SO=${1:-vspbfr_b200/libvsp_b200.so}
cuobjdump -sass "$SO" | c++filt | awk '
/Function :/ { fn=$0; sub(/.*Function : /, "", fn); gsub(/vsp::\(anonymous namespace\)::/, "", fn); gsub(/void /, "", fn);
               sub(/\(.*/, "", fn); names[fn]=1 }
/UTC[A-Z]*MMA/ { mma[fn]++ } /LDTM/ { ldtm[fn]++ } /UTMALDG/ { ldg[fn]++ } /UTMASTG/ { stg[fn]++ } /UTCBAR/ { bar[fn]++ }
/FFMA2|FMUL2|FADD2/ { f2[fn]++ } /\.2CTA/ { c2[fn]++ } /UBLKCP/ { blk[fn]++ } /LDG\.E[A-Z0-9.]*\.256/ { l256[fn]++ }
END { printf "%-8s %-6s %-8s %-8s %-7s %-6s %-7s %-7s %-7s %s\n", "UTC*MMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", ".2CTA", "UBLKCP", "LDG.256", "F*2", "kernel";
      for (f in names) printf "%-8d %-6d %-8d %-8d %-7d %-6d %-7d %-7d %-7d %s\n", mma[f], ldtm[f], ldg[f], stg[f], bar[f], c2[f], blk[f], l256[f], f2[f], f }' | (read h; echo "$h"; sort -k10)

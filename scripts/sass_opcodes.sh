#!/bin/bash
# SASS evidence: counts of the Blackwell-specific opcodes per kernel of the shipped library (B200_PROFILING.md "What proves a Blackwell-native kernel")
SO=${1:-gen_adversarial_b200/libga_b200.so}
cuobjdump -sass "$SO" | awk '
  /Function :/ { fn=$3 }
  /UTCHMMA|UTCQMMA|UTCIMMA|UTCOMMA/ { mma[fn]++ }
  /UTMALDG/ { ldg[fn]++ }
  /UTMASTG/ { stg[fn]++ }
  /UBLKCP/ { blk[fn]++ }
  /LDTM/ { ldtm[fn]++ }
  /UTCBAR/ { bar[fn]++ }
  /[^C]HMMA|HGMMA|QGMMA/ { legacy[fn]++ }
  END {
    printf "%-110s %8s %8s %8s %7s %6s %7s %7s\n", "kernel (mangled)", "UTC*MMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "UTCBAR", "legacy-MMA";
    for (f in mma) printf "%-110s %8d %8d %8d %7d %6d %7d %7d\n", substr(f,1,110), mma[f], ldg[f], stg[f], blk[f], ldtm[f], bar[f], legacy[f];
    for (f in ldg) if (!(f in mma)) printf "%-110s %8d %8d %8d %7d %6d %7d %7d\n", substr(f,1,110), 0, ldg[f], stg[f], blk[f], ldtm[f], bar[f], legacy[f];
  }' | sort

#!/bin/bash
# usage: tools/sass_census.sh > profiles/rNN_sass_census.txt
# Per kernel of the built objects: counts of the SASS mnemonics the design argues with (local memory STL/LDL, global
# reductions RED/ATOMG, FP64 tensor DMMA, bulk copies UBLKCP, cp.async LDGSTS, 256-bit loads / stores, FP64 FMA).
cd "$(dirname "$0")/../extendablefembase.jl_b200/build" || exit 1
for f in *.o; do
  cuobjdump -sass $f | awk -v F=${f%.o} '
    /Function :/ {name=$3; names[name]=1}
    /^[ \t]+\/\*[0-9a-f]+\*\// {
      if ($0 ~ /STL|LDL/) a[name,"STL/LDL"]++
      if ($0 ~ /RED\.|REDG|ATOMG/) a[name,"RED/ATOMG"]++
      if ($0 ~ /DMMA/) a[name,"DMMA"]++
      if ($0 ~ /UBLKCP/) a[name,"UBLKCP"]++
      if ($0 ~ /LDGSTS/) a[name,"LDGSTS"]++
      if ($0 ~ /LDG\.E\.(ENL2\.)?256/) a[name,"LDG.256"]++
      if ($0 ~ /STG\.E\.(ENL2\.)?256/) a[name,"STG.256"]++
      if ($0 ~ /DFMA/) a[name,"DFMA"]++
      if ($0 ~ /SYNCS/) a[name,"SYNCS(mbarrier)"]++
      n[name]++
    }
    END { for (k in names) { line = F " " k " instr=" n[k]; split("STL/LDL RED/ATOMG DMMA UBLKCP LDGSTS LDG.256 STG.256 DFMA SYNCS(mbarrier)", t, " "); for (i = 1; i <= 9; i++) if (a[k,t[i]]) line = line " " t[i] "=" a[k,t[i]]; print line } }'
done | sort | c++filt | grep -v "cub::\|thrust::"

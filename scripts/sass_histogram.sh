#!/bin/bash
# SASS opcode histogram of the shipped library (the evidence for which hardware paths the kernels use): per kernel the counts of
# the FP64 tensor-core (DMMA), bulk / async copy (UBLKCP, LDGSTS), mbarrier (SYNCS), 128-bit global / shared access opcodes.
so=${1:-limo_b200/libkba_b200.so}
echo "# SASS opcode histogram of \`$so\` (\`cuobjdump -sass\`, sm_100a)"
echo
echo "| kernel | instructions | DMMA | UBLKCP (cp.async.bulk) | LDGSTS (cp.async) | SYNCS (mbarrier) | USETMAXREG | LDG.E.128 | LDG.E.64 | STG.E.128 | STG.E.64 | LDS.128 | STS.128 | DFMA+DMUL+DADD | spills (STL/LDL) |"
echo "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"
cuobjdump -sass "$so" | awk '
/Function :/ { if (name != "") flush(); name=$3; n=0; delete c }
/^\s+\/\*[0-9a-f]+\*\// { n++; op=$2; if (op ~ /^@/) op=$3; sub(/;$/, "", op);
  if (op ~ /^DMMA/) c["dmma"]++; if (op ~ /^UBLKCP/) c["ublk"]++; if (op ~ /^LDGSTS/) c["ldgsts"]++; if (op ~ /^SYNCS/) c["syncs"]++;
  if (op ~ /^USETMAXREG/) c["maxreg"]++; if (op ~ /^LDG.*128/) c["ldg128"]++; if (op ~ /^LDG.*\.64/) c["ldg64"]++;
  if (op ~ /^STG.*128/) c["stg128"]++; if (op ~ /^STG.*\.64/) c["stg64"]++; if (op ~ /^LDS.*128/) c["lds128"]++; if (op ~ /^STS.*128/) c["sts128"]++;
  if (op ~ /^(DFMA|DMUL|DADD)/) c["dfma"]++; if (op ~ /^(STL|LDL)/) c["spill"]++ }
function flush() { printf("| %s | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d |\n", name, n, c["dmma"], c["ublk"], c["ldgsts"], c["syncs"], c["maxreg"], c["ldg128"], c["ldg64"], c["stg128"], c["stg64"], c["lds128"], c["sts128"], c["dfma"], c["spill"]) }
END { flush() }' | sed 's/_ZN3kba//' | sort

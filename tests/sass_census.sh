# SASS opcode census of the in-tree library (proof that the hot path is hand-written Blackwell code: tcgen05 = UTC*, tensor
# memory = LDTM / STTM, TMA = UTMA*, programmatic dependent launch = ACQBULK / griddepcontrol lowering).
#   bash tests/sass_census.sh > profiles/r02_sass_census.txt
SO=qwen3_asr_mlx_b200/lib/libqasr.so
echo "# SASS opcode census of $SO (sm_100a), round 2, final code"
echo "# command: bash tests/sass_census.sh   (cuobjdump -sass | mnemonic histogram; built by __graft_entry__.build(), $(nvcc --version | grep release | sed 's/.*release/nvcc release/'))"
echo
echo "## Blackwell-native and special-function instructions"
cuobjdump -sass $SO | grep -oE '\b(UTC[A-Z0-9_.]*|LDTM[.a-z0-9]*|STTM[.a-z0-9]*|UTMA[A-Z0-9_.]*|SYNCS[A-Z0-9_.]*|MUFU\.[A-Z0-9]*|HMMA[.A-Z0-9]*|REDG[.A-Za-z0-9]*|ACQBULK|DEPBAR[.A-Z0-9]*|ELECT|UCGABAR[_A-Z.]*|CCTL[.A-Z0-9]*|PREEXIT|ACQSHMINIT)' | sort | uniq -c | sort -rn
echo
echo "## programmatic dependent launch: griddepcontrol.wait -> ACQBULK, griddepcontrol.launch_dependents -> PREEXIT"
cuobjdump -sass $SO | grep -oE '\b(ACQBULK|PREEXIT)\b' | sort | uniq -c
echo
echo "## kernels in the library (entry points)"
cuobjdump -sass $SO | grep "Function :" | sed 's/.*Function : //' | c++filt | sed 's/(.*//' | sort | uniq -c | sort -rn
echo
echo "## shared libraries the .so links (no cuBLAS / cuDNN / NCCL)"
ldd $SO | awk '{print $1}'

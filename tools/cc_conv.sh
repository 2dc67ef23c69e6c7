#!/bin/bash
# compile stc_conv.cu alone and report the MMA issue path of one kernel (developer helper)
set -e
R=/root/repo/sentinel_tree_cover_b200
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -c $R/csrc/stc_conv.cu -o $R/build/stc_conv.o 2>&1 | grep -v deprec || true
echo "BRA.U.ANY count: $(cuobjdump -sass $R/build/stc_conv.o | grep -c 'BRA.U.ANY')"
F=${1:-_Z20conv3x3_umma2_kernelILi64ELi4ELi16ELi0ELb1EEv10ConvParamsiiiiii}
cuobjdump -sass -fun "$F" $R/build/stc_conv.o > /tmp/k.sass
grep -n "UTCHMMA" /tmp/k.sass | head -${2:-12}

#!/bin/bash
# Experiment builds: tools/build_variant.sh NAME FILE.cu "<extra nvcc flags>" [source override]
# -> build_variants/NAME/libb200force.so (FILE.cu recompiled with the flags, the other objects as built in csrc/).
# Select at run time with B200_LIB=build_variants/NAME/libb200force.so (mp-gadget_b200/__init__.py).
set -e
NAME=$1; FILE=$2; FLAGS=$3; SRC=${4:-mp-gadget_b200/csrc/$FILE}
D=build_variants/$NAME; mkdir -p $D
C=mp-gadget_b200/csrc
EXTRA=""; case $FILE in sph.cu|steploop.cu|domain_keys.cu|fof.cu) EXTRA="-fmad=false";; esac
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ -I$C $EXTRA $FLAGS -c $SRC -o $D/${FILE%.cu}.o
OBJS=""; for f in capi pm pm_fft pm_slab sharded tree_build tree_walk sph steploop domain_keys fof; do
  if [ "$f.cu" == "$FILE" ]; then OBJS="$OBJS $D/$f.o"; else OBJS="$OBJS $C/$f.o"; fi; done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $D/libb200force.so $OBJS -L/usr/local/cuda/lib64 -lcufft -ldl -Xlinker -rpath -Xlinker /usr/local/cuda/lib64
echo built $D/libb200force.so

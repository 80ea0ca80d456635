#!/bin/bash
# Build libssw from a SNAPSHOT of the sources (nvcc preprocesses the host pass minutes after the device pass: editing the
# tree while it runs gives inconsistent objects).  Usage: bash tools/build_tmp.sh <tag> [extra nvcc flags] -> /tmp/libssw_<tag>.so
TAG=$1; shift
D=/tmp/build_$TAG
rm -rf $D; mkdir -p $D/spread_spectrum_watermarking_b200 $D/include
cp -r spread_spectrum_watermarking_b200/csrc $D/spread_spectrum_watermarking_b200/
cp include/ssw.h $D/include/
cd $D/spread_spectrum_watermarking_b200/csrc && rm -f libssw.so
nvcc "$@" -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -shared -o /tmp/libssw_$TAG.so ssw_api.cu > /tmp/build_$TAG.log 2>&1
echo "nvcc rc=$?" >> /tmp/build_$TAG.log

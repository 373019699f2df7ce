#!/bin/bash
# A/B of two builds of the library on one box: adfwi_b200/csrc/libadfwi_b200.so (base) against adfwi_b200/csrc/lib_variant.so
# usage: bash tools/ab_variant.sh <script to run for each>      (e.g. tools/c2_quick.sh)
L=adfwi_b200/csrc
cp $L/libadfwi_b200.so /tmp/base.so
for round in 1 2; do
  echo "== base"; cp /tmp/base.so $L/libadfwi_b200.so; bash "$@"
  echo "== variant"; cp $L/lib_variant.so $L/libadfwi_b200.so; bash "$@"
done
cp /tmp/base.so $L/libadfwi_b200.so

#!/bin/bash
# Static SASS view of one kernel variant: total size and instructions per bit step of the posSlot tree
# (LDS at +0xb80 marks each step) -- no GPU needed.
V=${1:-3}
cuobjdump -sass /root/repo/lzma_b200/liblzgpu.so | awk "/Function : .*ILb0ELi${V}E/{f=1;next} /Function :/{f=0} f" | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed 's#/\*[0-9a-f]*\*/##g' | cut -c20-100 > /tmp/k.txt
echo "variant $V: $(wc -l < /tmp/k.txt) SASS instrs, BRA $(grep -c 'BRA' /tmp/k.txt), BSSY $(grep -c BSSY /tmp/k.txt)"
grep -n "LDS.U16 R[0-9]*, \[R[0-9]*+0xb8[02]\]" /tmp/k.txt | head -14 | awk -F: 'NR>1{printf "%d ", $1-p} {p=$1} END{print ""}'

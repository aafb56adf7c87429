#!/bin/bash
# usage: profiles/rf_eval.sh <lib.so or .o> <substring of kernel name>
f=$1; pat=$2
name=$(cuobjdump -elf $f 2>/dev/null | grep -o "_ZN[A-Za-z0-9_]*k_model_chisq[A-Za-z0-9_]*" | sort -u | grep "$pat" | grep -v _param | head -1)
cuobjdump -sass -fun "$name" $f > /tmp/rf_eval.sass 2>/dev/null
python $(dirname $0)/sass_rf_model.py /tmp/rf_eval.sass

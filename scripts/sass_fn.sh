#!/bin/bash
# usage: sass_fn.sh <all.sass> <mangled-name-substring>  -> prints the function's section
awk -v pat="$2" '/^\/\/-+ \.text\./{f=0} $0 ~ "^//-+ \\.text\\." && index($0,pat){f=1} f{print}' "$1"

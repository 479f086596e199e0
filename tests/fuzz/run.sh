#!/bin/sh
# Sanitizer fuzz of the native TIFF / PNG readers (tiff_host.h, png_host.h): 4,000 (TIFF) / 6,000 (PNG) random
# corruptions of every seed file -- byte flips in headers, tags and compressed data, truncations -- probed, read whole
# and read as a random region under AddressSanitizer + UBSan.  Any finding aborts with a report.
set -e
here=$(cd "$(dirname "$0")" && pwd)
work=${1:-/tmp/lars_fuzz}
mkdir -p "$work"
python "$here/make_seeds.py" "$work"
g++ -O1 -g -fsanitize=address,undefined -std=c++17 -pthread "$here/tiff_fuzz.cpp" -o "$work/tiff_fuzz" -ldl
g++ -O1 -g -fsanitize=address,undefined -std=c++17 -pthread "$here/png_fuzz.cpp" -o "$work/png_fuzz" -ldl
"$work/tiff_fuzz" "$work"/t*.tif
"$work/png_fuzz" "$work"/p*.png

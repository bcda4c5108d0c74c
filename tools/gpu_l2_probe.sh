for hint in 0 1; do for lag in 1 2 4 8 12; do
  ncu --metrics dram__bytes_read.sum --clock-control none -s 1 -c 1 --csv tools/bin/l2_probe 8 $lag 37 $hint 2>/dev/null | grep -E "dram__bytes_read|^chunk" | awk -F'","' '{ if (NF>3) print "   dram_read_MB", $(NF)/1e6; else print $0 }' | tr -d '"'
done; done
tools/bin/l2_probe 8 2 37 0; tools/bin/l2_probe 8 2 0 0

#!/bin/bash
# torchrun --no-python entry: rank 0 runs under ncu (NVLink + DRAM counters of the gather kernel only), the others plain
MODE=$1; OUT=$2
M="nvlrx__bytes.sum,nvltx__bytes.sum,nvlrx__bytes_data_user.sum,nvltx__bytes_data_user.sum,nvlrx__bytes_data_protocol.sum,nvltx__bytes_data_protocol.sum,nvlrx__bytes_packet_request.sum,nvlrx__bytes_packet_response.sum,nvltx__bytes_packet_request.sum,nvltx__bytes_packet_response.sum,gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"
rm -f /tmp/nvlink_ncu_done_$MODE
if [ "$LOCAL_RANK" = "0" ]; then
  exec ncu --metrics $M --clock-control none -k regex:dds_gather -s 4 -c 6 --csv --log-file $OUT python scripts/probes/nvlink_ncu.py $MODE /tmp/nvlink_ncu_done_$MODE
else
  exec python scripts/probes/nvlink_ncu.py $MODE /tmp/nvlink_ncu_done_$MODE
fi

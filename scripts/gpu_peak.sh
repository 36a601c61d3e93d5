for w in 4 8 32; do QK_PEAK_WARPS_PER_SM=$w python -c "
import sys; sys.path.insert(0,'qml-cutensornet_b200')
import qkmps; print('same-operand   warps/SM', $w, 'DMMA TFLOP/s', round(qkmps.dmma_peak(0, 40000),2))"; QK_PEAK_DISTINCT=1 QK_PEAK_WARPS_PER_SM=$w python -c "
import sys; sys.path.insert(0,'qml-cutensornet_b200')
import qkmps; print('distinct-operand warps/SM', $w, 'DMMA TFLOP/s', round(qkmps.dmma_peak(0, 40000),2))"; done

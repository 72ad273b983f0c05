/* [sum_short, diff_short, sum_long, diff_long] =
 *     DMX_frame_process(left, right, point_short, filter_coef, match_windowed, FFT_num, mtdWh, mtd_FFT_num, MTD_0_num)
 *
 * MEX gateway for one frame of MatlabProcess_xuzerui/CFAR_WangCai/DMX_SignalProcessing_main_xzr.m:332-353 (split + FIR /
 * circular matched filter), :414-426 (Hamming MTD zero-padded to mtd_FFT_num, |L|+|R|, |R|-|L|) and :462-465 (zero-Doppler
 * blanking of the sums).  The script has no function for this block, so there is no M file to shadow: a maintainer replaces
 * the lines above by this one call, passing the variables the script already holds
 *     (echoData_Frame_Left, echoData_Frame_Right, point_short, filter_coef, matchWaveform2.*mfWh.', FFT_num, mtdWh,
 *      mtd_FFT_num, MTD_0_num)
 * and gets echo_MTD_sum_short / diff_short / sum_long / diff_long back for executeCFAR and motionParaMeasure.           */
#include "rb200_mex_common.h"

extern "C" void mexFunction(int nlhs, mxArray* plhs[], int nrhs, const mxArray* prhs[]) {
    rb_require(nrhs == 9, "radar_b200:dmx:nargin", "DMX_frame_process: expected 9 inputs");
    rb_require(nlhs <= 4, "radar_b200:dmx:nargout", "DMX_frame_process: at most four outputs");
    for (int i = 0; i < 2; ++i) rb_require_real_or_complex_double(prhs[i], "radar_b200:dmx:type");
    rb_require_real_or_complex_double(prhs[3], "radar_b200:dmx:type");
    rb_require_real_or_complex_double(prhs[4], "radar_b200:dmx:type");
    rb_require_real_or_complex_double(prhs[6], "radar_b200:dmx:type");
    const int P = (int)mxGetM(prhs[0]), n_range = (int)mxGetN(prhs[0]);
    rb_require((int)mxGetM(prhs[1]) == P && (int)mxGetN(prhs[1]) == n_range, "radar_b200:dmx:dimensionMismatch",
               "DMX_frame_process: the two beams differ in size");
    const int n_short = (int)rb_scalar(prhs[2], "radar_b200:dmx:type");
    const int n_fir = (int)mxGetNumberOfElements(prhs[3]);
    const int n_mf = (int)mxGetNumberOfElements(prhs[4]);
    const int fft_num = (int)rb_scalar(prhs[5], "radar_b200:dmx:type");
    rb_require((int)mxGetNumberOfElements(prhs[6]) == P, "radar_b200:dmx:dimensionMismatch",
               "DMX_frame_process: mtdWh must have prtNum entries (Arrays have incompatible sizes)");
    const int mtd_fft = (int)rb_scalar(prhs[7], "radar_b200:dmx:type");
    const int n0 = (int)rb_scalar(prhs[8], "radar_b200:dmx:type");
    rb_require(n_short >= 0 && n_short < n_range && fft_num >= 1 && mtd_fft >= 1, "radar_b200:dmx:indexOutOfRange",
               "DMX_frame_process: point_short / FFT sizes out of range");
    mxArray* out[4];
    out[0] = mxCreateDoubleMatrix(mtd_fft, n_short, mxREAL);
    out[1] = mxCreateDoubleMatrix(mtd_fft, n_short, mxREAL);
    out[2] = mxCreateDoubleMatrix(mtd_fft, fft_num, mxREAL);
    out[3] = mxCreateDoubleMatrix(mtd_fft, fft_num, mxREAL);
    const int st = rb200_dmx_process_z(rb_context(), mxGetPr(prhs[0]), mxGetPi(prhs[0]), mxGetPr(prhs[1]), mxGetPi(prhs[1]), P, n_range,
                                       n_short, mxGetPr(prhs[3]), n_fir, mxGetPr(prhs[4]), mxGetPi(prhs[4]), n_mf, fft_num,
                                       mxGetPr(prhs[6]), mtd_fft, n0, n_short ? mxGetPr(out[0]) : NULL, n_short ? mxGetPr(out[1]) : NULL,
                                       mxGetPr(out[2]), mxGetPr(out[3]));
    const int wanted = nlhs < 1 ? 1 : nlhs;
    for (int i = 0; i < 4; ++i) {
        if (i < wanted) plhs[i] = out[i];
        else mxDestroyArray(out[i]);
    }
    rb_check(st, "dmx");
}

// gas_filter.cuh — upstream AudioFilterSW::prepare_coefficients for the device, shared by the prologue and the per-call
// (single voice) entry points.  Include only from translation units compiled with -fmad=false: the coefficients are double
// arithmetic narrowed to float exactly where upstream narrows.
#pragma once

#include "gas_internal.h"

// upstream AudioFilterSW::prepare_coefficients: double arithmetic, every coefficient
// narrowed to float on store and once more after the division by a0; feedback terms stored negated.
static __device__ void prepare_coefficients(int mode, float cutoff, float resonance, float gain, int stages, float sampling_rate, float out[5]) {
	int sr_limit = (int)((sampling_rate / 2) + 512);
	double final_cutoff = (cutoff > sr_limit) ? (double)sr_limit : (double)cutoff;
	if (final_cutoff < 1) {
		final_cutoff = 1;
	}
	const double TAU = 6.2831853071795864769252867666;
	double omega = TAU * final_cutoff / (double)sampling_rate;
	double sin_v = sin(omega);
	double cos_v = cos(omega);
	double Q = resonance;
	if (Q <= 0.0) {
		Q = 0.0001;
	}
	if (mode == GAS_FILTER_BANDPASS) {
		Q *= 2.0;
	} else if (mode == GAS_FILTER_PEAK) {
		Q *= 3.0;
	}
	double tmpgain = gain;
	if (tmpgain < 0.001) {
		tmpgain = 0.001;
	}
	if (stages > 1) {
		Q = (Q > 1.0 ? pow(Q, 1.0 / stages) : Q);
		tmpgain = pow(tmpgain, 1.0 / (stages + 1));
	}
	double alpha = sin_v / (2 * Q);
	double a0 = 1.0 + alpha;
	float b0 = 0.f, b1 = 0.f, b2 = 0.f, a1 = 0.f, a2 = 0.f;
	switch (mode) {
		case GAS_FILTER_LOWPASS:
			b0 = (float)((1.0 - cos_v) / 2.0);
			b1 = (float)(1.0 - cos_v);
			b2 = (float)((1.0 - cos_v) / 2.0);
			a1 = (float)(-2.0 * cos_v);
			a2 = (float)(1.0 - alpha);
			break;
		case GAS_FILTER_HIGHPASS:
			b0 = (float)((1.0 + cos_v) / 2.0);
			b1 = (float)(-(1.0 + cos_v));
			b2 = (float)((1.0 + cos_v) / 2.0);
			a1 = (float)(-2.0 * cos_v);
			a2 = (float)(1.0 - alpha);
			break;
		case GAS_FILTER_BANDPASS:
			b0 = (float)(alpha * sqrt(Q + 1));
			b1 = 0.f;
			b2 = (float)(-alpha * sqrt(Q + 1));
			a1 = (float)(-2.0 * cos_v);
			a2 = (float)(1.0 - alpha);
			break;
		case GAS_FILTER_NOTCH:
			b0 = 1.f;
			b1 = (float)(-2.0 * cos_v);
			b2 = 1.f;
			a1 = (float)(-2.0 * cos_v);
			a2 = (float)(1.0 - alpha);
			break;
		case GAS_FILTER_PEAK:
			b0 = (float)(1.0 + alpha * tmpgain);
			b1 = (float)(-2.0 * cos_v);
			b2 = (float)(1.0 - alpha * tmpgain);
			a1 = (float)(-2 * cos_v);
			a2 = (float)(1 - alpha / tmpgain);
			break;
		case GAS_FILTER_BANDLIMIT: {
			double hicutoff = resonance;
			double centercutoff = ((double)cutoff + (double)resonance) / 2.0;
			double bandwidth = (log(centercutoff) - log(hicutoff)) / log(2.0);
			omega = TAU * centercutoff / (double)sampling_rate;
			alpha = sin(omega) * sinh(log(2.0) / 2 * bandwidth * omega / sin(omega));
			a0 = 1 + alpha;
			b0 = (float)alpha;
			b1 = 0.f;
			b2 = (float)-alpha;
			a1 = (float)(-2 * cos(omega));
			a2 = (float)(1 - alpha);
		} break;
		case GAS_FILTER_LOWSHELF: {
			double tmpq = sqrt(Q);
			if (tmpq <= 0) {
				tmpq = 0.001;
			}
			double beta = sqrt(tmpgain) / tmpq;
			a0 = (tmpgain + 1.0) + (tmpgain - 1.0) * cos_v + beta * sin_v;
			b0 = (float)(tmpgain * ((tmpgain + 1.0) - (tmpgain - 1.0) * cos_v + beta * sin_v));
			b1 = (float)(2.0 * tmpgain * ((tmpgain - 1.0) - (tmpgain + 1.0) * cos_v));
			b2 = (float)(tmpgain * ((tmpgain + 1.0) - (tmpgain - 1.0) * cos_v - beta * sin_v));
			a1 = (float)(-2.0 * ((tmpgain - 1.0) + (tmpgain + 1.0) * cos_v));
			a2 = (float)((tmpgain + 1.0) + (tmpgain - 1.0) * cos_v - beta * sin_v);
		} break;
		default: { // HIGHSHELF
			double tmpq = sqrt(Q);
			if (tmpq <= 0) {
				tmpq = 0.001;
			}
			double beta = sqrt(tmpgain) / tmpq;
			a0 = (tmpgain + 1.0) - (tmpgain - 1.0) * cos_v + beta * sin_v;
			b0 = (float)(tmpgain * ((tmpgain + 1.0) + (tmpgain - 1.0) * cos_v + beta * sin_v));
			b1 = (float)(-2.0 * tmpgain * ((tmpgain - 1.0) + (tmpgain + 1.0) * cos_v));
			b2 = (float)(tmpgain * ((tmpgain + 1.0) + (tmpgain - 1.0) * cos_v - beta * sin_v));
			a1 = (float)(2.0 * ((tmpgain - 1.0) - (tmpgain + 1.0) * cos_v));
			a2 = (float)((tmpgain + 1.0) - (tmpgain - 1.0) * cos_v - beta * sin_v);
		} break;
	}
	out[0] = (float)((double)b0 / a0);
	out[1] = (float)((double)b1 / a0);
	out[2] = (float)((double)b2 / a0);
	out[3] = (float)((double)a1 / (0.0 - a0));
	out[4] = (float)((double)a2 / (0.0 - a0));
}


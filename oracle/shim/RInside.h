// oracle/shim/RInside.h -- stand-in for RInside/Rcpp, which are absent from this image.
// Only what the reference hot path touches is provided:
//   * Tools::init_r() (Tools/UtilFunctions.cpp:84-95) needs RInside::instancePtr/parseEvalQ;
//   * PoissonTargetEstimator.cpp:88 needs Rcpp::ppois(IntegerVector, lambda, lower_tail)[0].
// ppois restates R's documented definition: ppois(q, l, lower.tail=FALSE) = P[X > q], X ~ Poisson(l)
// (third-party arithmetic: R >= 3.2.2 `ppois`, unpinned by the reference; results feed a threshold compare only).
#pragma once
#include <cmath>
#include <string>
#include <vector>

class RInside
{
public:
	RInside(int, const char *const *) { self() = this; }
	static RInside *instancePtr() { return self(); }
	void parseEvalQ(const std::string &) {}
private:
	static RInside *&self() { static RInside *p = nullptr; return p; }
};

namespace Rcpp
{
	struct IntegerVector
	{
		std::vector<long> v;
		static IntegerVector create(long x) { IntegerVector r; r.v.push_back(x); return r; }
	};

	namespace shim_detail
	{
		inline double log_pmf(long k, double lambda) { return -lambda + k * std::log(lambda) - std::lgamma(double(k) + 1.0); }

		// P[X <= x]: terms summed from the largest one outwards (stable for any lambda).
		inline double lower(long x, double lambda)
		{
			if (x < 0) return 0;
			if (lambda <= 0) return 1;
			double s = 0;
			for (long k = x; k >= 0; --k)
			{
				double t = std::exp(log_pmf(k, lambda));
				s += t;
				if (double(k) < lambda && t < s * 1e-17) break;
			}
			return s > 1 ? 1 : s;
		}

		// P[X > x]
		inline double upper(long x, double lambda)
		{
			if (x < 0) return 1;
			if (lambda <= 0) return 0;
			if (double(x + 1) < lambda) return 1 - lower(x, lambda);
			double s = 0;
			for (long k = x + 1;; ++k)
			{
				double t = std::exp(log_pmf(k, lambda));
				s += t;
				if (t <= s * 1e-17 || k > x + 100000) break;
			}
			return s > 1 ? 1 : s;
		}
	}

	inline std::vector<double> ppois(const IntegerVector &q, double lambda, bool lower_tail)
	{
		std::vector<double> res;
		for (long x : q.v) res.push_back(lower_tail ? shim_detail::lower(x, lambda) : shim_detail::upper(x, lambda));
		return res;
	}
}

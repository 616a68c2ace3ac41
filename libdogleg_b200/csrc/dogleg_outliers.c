/* dogleg_outliers.c -- outlier / confidence helpers (reference dogleg.c:1826-3149,
 * API dogleg.h:333-392). SURVEY.md 8f-1: "next" row, post-solve consumers of the
 * factorization through multi-RHS solves. Placeholder until that row is built:
 * the symbols exist so that programs link, and fail loudly. */
#include <stdio.h>
#include "dogleg.h"
#define SAY(fmt, ...) fprintf(stderr, "libdogleg at %s:%d: " fmt "\n", __FILE__, __LINE__, ## __VA_ARGS__)

bool dogleg_getOutliernessFactors(double* factors, double* scale, int featureSize, int Nfeatures,
                                  int NoutlierFeatures, dogleg_operatingPoint_t* point, dogleg_solverContext_t* ctx)
{
  (void)factors; (void)scale; (void)featureSize; (void)Nfeatures; (void)NoutlierFeatures; (void)point; (void)ctx;
  SAY("dogleg_getOutliernessFactors() is not available in this build yet");
  return false;
}
bool dogleg_markOutliers(struct dogleg_outliers_t* markedOutliers, double* scale, int* Noutliers,
                         double (getConfidence)(int i_feature_exclude), int featureSize, int Nfeatures,
                         dogleg_operatingPoint_t* point, dogleg_solverContext_t* ctx)
{
  (void)markedOutliers; (void)scale; (void)Noutliers; (void)getConfidence; (void)featureSize; (void)Nfeatures; (void)point; (void)ctx;
  SAY("dogleg_markOutliers() is not available in this build yet");
  return false;
}
void dogleg_reportOutliers(double (getConfidence)(int i_feature_exclude), double* scale, int featureSize,
                           int Nfeatures, int Noutliers, dogleg_operatingPoint_t* point, dogleg_solverContext_t* ctx)
{
  (void)getConfidence; (void)scale; (void)featureSize; (void)Nfeatures; (void)Noutliers; (void)point; (void)ctx;
  SAY("dogleg_reportOutliers() is not available in this build yet");
}
double dogleg_getOutliernessTrace_newFeature_sparse(const double* JqueryFeature, int istateActive, int NstateActive,
                                                    int featureSize, int NoutlierFeatures,
                                                    dogleg_operatingPoint_t* point, dogleg_solverContext_t* ctx)
{
  (void)JqueryFeature; (void)istateActive; (void)NstateActive; (void)featureSize; (void)NoutlierFeatures; (void)point; (void)ctx;
  SAY("dogleg_getOutliernessTrace_newFeature_sparse() is not available in this build yet");
  return -1.0;
}

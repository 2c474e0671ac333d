// ref_probe.cpp — TEST INFRASTRUCTURE.  Compiled against the reference headers and linked to the
// reference library (oracle/_ref/lib_lightgbm.so).  Reads the *binned* training matrix and the
// per-feature layout metadata out of a reference DatasetHandle so that the oracle, the CUDA path
// and the reference itself can be run on IDENTICAL binned input (SURVEY.md §8c "identical binned
// input").  A DatasetHandle is a LightGBM::Dataset* (reference src/c_api.cpp:1306-1330).
#include <LightGBM/bin.h>
#include <LightGBM/dataset.h>

#include <cstdint>
#include <memory>

using LightGBM::BinIterator;
using LightGBM::BinMapper;
using LightGBM::Dataset;

extern "C" {

int RefProbe_Dims(void* handle, int32_t* num_data, int32_t* num_columns, int32_t* num_features,
                  int32_t* num_total_features) {
  const Dataset* d = reinterpret_cast<const Dataset*>(handle);
  *num_data = d->num_data();
  *num_columns = d->num_feature_groups();
  *num_features = d->num_features();
  *num_total_features = d->num_total_features();
  return 0;
}

// Per inner feature: column (group), first stored value, num_bin, most_freq_bin, default_bin,
// missing type, real (original) feature index, #features in its group, bin type (0 numerical).
int RefProbe_Layout(void* handle, int32_t* feat_column, int32_t* feat_lo, int32_t* feat_num_bin,
                    int32_t* feat_mfb, int32_t* feat_default_bin, int32_t* feat_missing,
                    int32_t* feat_real_index, int32_t* feat_in_group, int32_t* feat_bin_type) {
  const Dataset* d = reinterpret_cast<const Dataset*>(handle);
  const int F = d->num_features();
  std::vector<int> group_size(d->num_feature_groups(), 0);
  for (int f = 0; f < F; ++f) group_size[d->Feature2Group(f)]++;
  for (int f = 0; f < F; ++f) {
    const BinMapper* bm = d->FeatureBinMapper(f);
    const int g = d->Feature2Group(f);
    if (d->IsMultiGroup(g)) return -2;  // multi-val (row-wise sparse) groups are outside the contract
    feat_column[f] = g;
    feat_lo[f] = static_cast<int32_t>(d->feature_min_bin(f));
    feat_num_bin[f] = bm->num_bin();
    feat_mfb[f] = static_cast<int32_t>(bm->GetMostFreqBin());
    feat_default_bin[f] = static_cast<int32_t>(bm->GetDefaultBin());
    feat_missing[f] = static_cast<int32_t>(bm->missing_type());
    feat_real_index[f] = d->RealFeatureIndex(f);
    feat_in_group[f] = group_size[g];
    feat_bin_type[f] = bm->bin_type() == LightGBM::BinType::NumericalBin ? 0 : 1;
  }
  return 0;
}

// Row-major [num_data x num_columns] stored group values (what FeatureGroup::PushData wrote).
int RefProbe_Bins(void* handle, uint8_t* out) {
  const Dataset* d = reinterpret_cast<const Dataset*>(handle);
  const int C = d->num_feature_groups();
  const int64_t N = d->num_data();
  for (int g = 0; g < C; ++g) {
    if (d->FeatureGroupNumBin(g) > 256) return -3;
    std::unique_ptr<BinIterator> it(d->FeatureGroupIterator(g));
    if (!it) return -2;
    it->Reset(0);
    for (int64_t i = 0; i < N; ++i) out[i * C + g] = static_cast<uint8_t>(it->RawGet(static_cast<int>(i)));
  }
  return 0;
}

// bin -> upper bound (Dataset::RealThreshold, dataset.h:853-857) for every bin of feature f.
int RefProbe_BinUpperBounds(void* handle, int f, double* out) {
  const Dataset* d = reinterpret_cast<const Dataset*>(handle);
  const BinMapper* bm = d->FeatureBinMapper(f);
  for (int b = 0; b < bm->num_bin(); ++b) out[b] = bm->BinToValue(b);
  return 0;
}

}  // extern "C"

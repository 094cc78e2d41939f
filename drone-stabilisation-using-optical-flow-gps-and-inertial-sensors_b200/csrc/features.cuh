// features.cuh -- shared declarations of the Shi-Tomasi stage.
#pragma once
#include "common.cuh"

struct FeatImageState {
    unsigned int max_key;     // order-preserving key of the running max of lambda_min (0 = none yet)
    unsigned int n_cand;      // candidates appended by kernel A
    int n_out;                // corners written by kernel B
    int overflow;             // candidate buffer too small
};

// Device-pointer core: n_images images of identical geometry, image i at img + i*istride; corners of
// image i at xy_out + i*xy_stride (floats), at most out_cap each. *state_out = per-image state array
// (device) valid until the next call on this context.
int ofb_features_device(ofb_ctx* ctx, const uint8_t* img, int w, int h, int pitch, size_t istride, int n_images,
                        const uint8_t* mask, int mpitch, size_t mstride, int max_corners, double quality,
                        double min_distance, int block_size, unsigned int cand_cap, float* xy_out, size_t xy_stride,
                        int out_cap, FeatImageState** state_out);

int ofb_features_scratch(ofb_ctx* ctx, int w, int h, double min_distance, int n_images, FeatImageState** st_out,
                         int** grid_out, size_t* cell_stride_out);

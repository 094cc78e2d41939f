// host_check.cu -- runs the __host__ __device__ pieces of the product (Philox, Box-Muller, the
// per-trial Monte-Carlo arithmetic, the 3x3 eigen/solve helpers) on the CPU so that they can be
// compared with the oracle without a GPU (tests/test_host_math.py). Test infrastructure only.
//
//   host_check <in.bin> <out.bin>
// in : ofb_mc_step | uint64 seed | uint32 step | uint32 ntrials | double pos[2N] | double flow[2N]
// out: for each precision p in {fp32, fp64}: ntrials x (v[3], R) doubles;
//      then 4*ntrials normals (fp64 path) of draw block 3 for the same trials;
//      then jacobi: eigenvalues[3] + eigenvectors[9] + newton lambda_min of the trial-0 normal matrix
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../drone-stabilisation-using-optical-flow-gps-and-inertial-sensors_b200/csrc/montecarlo.cu"

void ofb_set_error(const char*, ...) {}   // api.cu is not linked into the harness

int main(int argc, char** argv)
{
    if (argc < 3) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 3;
    ofb_mc_step st;
    uint64_t seed; uint32_t step, ntr;
    if (fread(&st, sizeof(st), 1, f) != 1) return 4;
    if (fread(&seed, 8, 1, f) != 1 || fread(&step, 4, 1, f) != 1 || fread(&ntr, 4, 1, f) != 1) return 4;
    int N = st.n_points;
    std::vector<double> pos(2 * N), flow(2 * N);
    if (fread(pos.data(), 8, 2 * N, f) != (size_t)(2 * N) || fread(flow.data(), 8, 2 * N, f) != (size_t)(2 * N)) return 4;
    fclose(f);
    uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    FILE* o = fopen(argv[2], "wb");
    {
        McParams<float> P; mc_load_params<float>(st, P);
        std::vector<float> p(2 * N), fl(2 * N);
        for (int i = 0; i < 2 * N; ++i) { p[i] = (float)pos[i]; fl[i] = (float)flow[i]; }
        std::vector<float> pc(2 * N);
        for (int j = 0; j < N; ++j) mc_point_consts<float>(P, p.data(), j, pc.data());
        for (uint32_t t = 0; t < ntr; ++t) {
            double v[4];
            mc_trial<float>(P, p.data(), fl.data(), pc.data(), t, key, step, v, v[3]);
            fwrite(v, 8, 4, o);
        }
    }
    {
        McParams<double> P; mc_load_params<double>(st, P);
        std::vector<double> pc(2 * N);
        for (int j = 0; j < N; ++j) mc_point_consts<double>(P, pos.data(), j, pc.data());
        for (uint32_t t = 0; t < ntr; ++t) {
            double v[4];
            mc_trial<double>(P, pos.data(), flow.data(), pc.data(), t, key, step, v, v[3]);
            fwrite(v, 8, 4, o);
        }
    }
    for (uint32_t t = 0; t < ntr; ++t) {
        double z[4];
        mc_normals4<double>(t, 0u, 3u, step, key, z[0], z[1], z[2], z[3]);
        fwrite(z, 8, 4, o);
    }
    {
        // a fixed SPD test matrix built from the points
        double M[6] = {0, 0, 0, 0, 0, 0};
        for (int j = 0; j < N; ++j) {
            double x = pos[2 * j], y = pos[2 * j + 1];
            M[0] += y * y + 1; M[1] -= x * y; M[2] -= x; M[3] += x * x + 1; M[4] -= y; M[5] += x * x + y * y;
        }
        double ev[3], q[3][3];
        ofb_jacobi3(M, ev, q);
        double lmin = ofb_min_eig_sym3(M);
        fwrite(M, 8, 6, o); fwrite(ev, 8, 3, o); fwrite(q, 8, 9, o); fwrite(&lmin, 8, 1, o);
    }
    fclose(o);
    return 0;
}

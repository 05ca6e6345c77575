/* A consumer of the tebscat C ABI with no Python and no CUDA headers at run time: loads a plan file written by
 * `python -m tebscat.export_plan`, transforms a batch of host signals and writes the coefficients.
 *
 *   gcc -O2 -I include examples/c_consumer.c -o c_consumer -L vae-teb_b200/tebscat -ltebscat -Wl,-rpath,vae-teb_b200/tebscat -lm
 *   ./c_consumer plan.tebplan signals.f32 coefficients.f32      (signals: B x N float32, row-major)
 *
 * This is the call the dataset builder makes per record (hdf5_dataset/create_hdf5_dataset.py:418-441:
 * .to(device) ... st_model(...) ... .cpu().numpy()), from any language with a C FFI. */
#include <stdio.h>
#include <stdlib.h>

#include "tebscat.h"

int main(int argc, char** argv) {
    if (argc != 4) {
        fprintf(stderr, "usage: %s plan.tebplan signals.f32 coefficients.f32\n", argv[0]);
        return 2;
    }
    tebscat_plan* plan = NULL;
    if (tebscat_plan_load(argv[1], 0, &plan) != TEBSCAT_OK) {
        fprintf(stderr, "tebscat_plan_load: %s\n", tebscat_last_error());
        return 1;
    }
    tebscat_plan_desc d;
    tebscat_plan_get_desc(plan, &d);
    FILE* f = fopen(argv[2], "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", argv[2]); return 1; }
    fseek(f, 0, SEEK_END);
    const long bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    const long B = bytes / (long)(sizeof(float) * d.N);
    if (B < 1 || B * (long)sizeof(float) * d.N != bytes) { fprintf(stderr, "%s is not B x %d float32\n", argv[2], d.N); return 1; }
    float* x = (float*)malloc((size_t)bytes);
    float* S = (float*)malloc((size_t)B * d.n_paths * d.n_out * sizeof(float));
    if (!x || !S || fread(x, 1, (size_t)bytes, f) != (size_t)bytes) { fprintf(stderr, "read failed\n"); return 1; }
    fclose(f);
    if (tebscat_scat1d_forward_host(plan, x, B, S) != TEBSCAT_OK) {
        fprintf(stderr, "tebscat_scat1d_forward_host: %s\n", tebscat_last_error());
        return 1;
    }
    f = fopen(argv[3], "wb");
    if (!f || fwrite(S, sizeof(float), (size_t)B * d.n_paths * d.n_out, f) != (size_t)B * d.n_paths * d.n_out) {
        fprintf(stderr, "write failed\n");
        return 1;
    }
    fclose(f);
    printf("%ld signals of %d samples -> %d channels x %d samples each\n", B, d.N, d.n_paths, d.n_out);
    tebscat_plan_destroy(plan);
    free(x);
    free(S);
    return 0;
}

/*
 * llz_resample_cuda -- command-line harness over libllzfilter_cuda with the flags and the output-length
 * semantics of the reference's example/llz_resample (main.c:22-130, llz_parseopt.c:164-297):
 *
 *   llz_resample_cuda -i in.wav -o out.wav [-t 0|1|2] [-u up] [-d down] [-g gain] [-w] [-c]
 *
 *   -w  whole-file mode: read all samples, zero-pad to the same frame count, and run them through the handle in
 *       ONE call of the batched entry point (llz_cuda_resample_bank_run_host) instead of one llz_resample call
 *       per frame.  The bytes written are identical (a call of any length continues the stream exactly like
 *       the frame loop); it only removes the per-frame PCIe round trips.
 *
 *   -c  channels mode: honour fmt.channels (libllzaudio/llz_wavfmt.h:22-38).  The reference filters a multi-channel file as
 *       ONE interleaved mono stream (quirk R7, main.c:60-62), which mixes the channels; with -c every channel is resampled
 *       on its own (a bank of fmt.channels channels) straight from the interleaved frames -- the de-interleave is fused
 *       into the kernel's load stage (llz_cuda_resample_bank_run_pcm_host) -- and the frames are written interleaved
 *       again.  Same frame count rule per channel as the default mode.  The default stays R7-compatible.
 *
 *   -t 0 decimate by -d, 1 interpolate by -u, 2 (default) resample by -u/-d; defaults 160/147, gain 1.
 *   Giving only -u sets down = 1, only -d sets up = 1 (llz_parseopt.c:137-141).  Window: BLACKMAN (main.c:67-75).
 *
 * The input must be a 16-bit PCM WAV; samples are read from byte 44 on (main.c:62) and treated as one mono
 * stream whatever the channel count (quirk R7).  The last partial frame is zero-padded, and a file that is
 * an exact number of frames still gets one extra all-zero frame (main.c:96-99), so
 * frames_out = floor(data_bytes / frame_bytes) + 1.  Every sample is computed on the GPU.
 */
#include <getopt.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "llz_cuda.h"

typedef struct {
    uint16_t format, channels, bytes_per_sample;
    uint32_t samplerate;
} wav_info_t;

static uint32_t le32(const unsigned char *p) { return p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint16_t le16(const unsigned char *p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static void put32(unsigned char *p, uint32_t v) { p[0] = v; p[1] = v >> 8; p[2] = v >> 16; p[3] = v >> 24; }
static void put16(unsigned char *p, uint16_t v) { p[0] = (unsigned char)v; p[1] = (unsigned char)(v >> 8); }

/* walk the RIFF chunks up to "fmt " (the reference does the same and then ignores where "data" starts) */
static int wav_read_info(FILE *fp, wav_info_t *w)
{
    unsigned char b[16];
    if (fread(b, 1, 12, fp) != 12 || memcmp(b, "RIFF", 4) || memcmp(b + 8, "WAVE", 4)) {
        fprintf(stderr, "file is not WAVE format!\n");
        return -1;
    }
    for (;;) {
        if (fread(b, 1, 8, fp) != 8) { fprintf(stderr, "no fmt chunk\n"); return -1; }
        uint32_t size = le32(b + 4);
        if (!memcmp(b, "fmt ", 4)) {
            unsigned char f[16];
            if (size < 16 || fread(f, 1, 16, fp) != 16) { fprintf(stderr, "short fmt chunk\n"); return -1; }
            w->format = le16(f);
            w->channels = le16(f + 2);
            w->samplerate = le32(f + 4);
            w->bytes_per_sample = (uint16_t)((le16(f + 14) + 7) / 8);
            return 0;
        }
        fseek(fp, (long)size, SEEK_CUR);
    }
}

/* canonical 44-byte PCM header (what llz_wavfmt_writeheader emits, llz_wavfmt.c:185-213) */
static void wav_write_header(FILE *fp, const wav_info_t *w, uint32_t data_bytes)
{
    unsigned char h[44];
    uint16_t align = (uint16_t)(w->channels * w->bytes_per_sample);
    memcpy(h, "RIFF", 4);            put32(h + 4, data_bytes + 36);
    memcpy(h + 8, "WAVEfmt ", 8);    put32(h + 16, 16);
    put16(h + 20, w->format);        put16(h + 22, w->channels);
    put32(h + 24, w->samplerate);    put32(h + 28, w->samplerate * align);
    put16(h + 32, align);            put16(h + 34, (uint16_t)(w->bytes_per_sample * 8));
    memcpy(h + 36, "data", 4);       put32(h + 40, data_bytes);
    fseek(fp, 0, SEEK_SET);
    fwrite(h, 1, 44, fp);
}

int main(int argc, char **argv)
{
    const char *in = NULL, *out = NULL;
    int type = 2, up = 160, down = 147, got_up = 0, got_down = 0, quiet = 0, whole = 0, chmode = 0;
    double gain = 1.0;
    static struct option lopts[] = {{"help", 0, 0, 'h'}, {"input", 1, 0, 'i'}, {"output", 1, 0, 'o'}, {"type", 1, 0, 't'},
                                    {"down", 1, 0, 'd'}, {"up", 1, 0, 'u'}, {"gain", 1, 0, 'g'}, {"quiet", 0, 0, 'q'}, {"whole", 0, 0, 'w'}, {"channels", 0, 0, 'c'}, {0, 0, 0, 0}};
    int c;
    while ((c = getopt_long(argc, argv, "hqwci:o:t:d:u:g:", lopts, NULL)) != -1) {
        switch (c) {
        case 'i': in = optarg; break;
        case 'o': out = optarg; break;
        case 't': type = atoi(optarg); break;
        case 'd': down = atoi(optarg); got_down = 1; break;
        case 'u': up = atoi(optarg); got_up = 1; break;
        case 'g': gain = atof(optarg); break;
        case 'q': quiet = 1; break;
        case 'w': whole = 1; break;
        case 'c': chmode = 1; break;
        default:
            fprintf(stderr, "usage: %s -i in.wav -o out.wav [-t 0|1|2] [-u up] [-d down] [-g gain]\n", argv[0]);
            return c == 'h' ? 0 : -1;
        }
    }
    if (!in || !out) { fprintf(stderr, "FAIL: input and output file should input\n"); return -1; }
    if (got_up && !got_down) down = 1;
    if (!got_up && got_down) up = 1;
    if (type < 0 || type > 2 || up < 1 || down < 1) { fprintf(stderr, "FAIL: bad type or factor\n"); return -1; }
    double ratio = (double)up / down;
    if (ratio > LLZ_RATIO_MAX || 1. / ratio > LLZ_RATIO_MAX) {
        fprintf(stderr, "FAIL: ratio not support, you can use cascade method to implement \n");
        return -1;
    }

    FILE *fo = fopen(out, "w+b"), *fi = fopen(in, "rb");
    if (!fo || !fi) { fprintf(stderr, "%s file can not be opened\n", fo ? "input" : "output"); return -1; }
    wav_info_t w;
    if (wav_read_info(fi, &w) != 0) return -1;
    if (w.format != 1) { fprintf(stderr, "error! unsupported WAVE file format.\n"); return -1; }
    fseek(fi, 44, SEEK_SET);

    if (chmode) {
        /* one bank channel per WAV channel, interleaved frames in and out */
        if (w.bytes_per_sample != 2 || w.channels < 1) { fprintf(stderr, "error! -c needs 16-bit PCM\n"); return -1; }
        const int C = w.channels;
        unsigned long hb = type == 0 ? llz_cuda_decimate_bank_init(down, gain, BLACKMAN, C, LLZ_CUDA_ACC_F64)
                         : type == 1 ? llz_cuda_interp_bank_init(up, gain, BLACKMAN, C, LLZ_CUDA_ACC_F64)
                                     : llz_cuda_resample_bank_init(up, down, gain, BLACKMAN, 0, C, LLZ_CUDA_ACC_F64);
        if (hb == (unsigned long)-1) { fprintf(stderr, "init failed: %s\n", llz_cuda_last_error()); return -1; }
        llz_cuda_resample_info_t bi_;
        llz_cuda_resample_bank_info(hb, &bi_);
        long pos = ftell(fi);
        fseek(fi, 0, SEEK_END);
        long data_bytes = ftell(fi) - pos;
        fseek(fi, pos, SEEK_SET);
        if (data_bytes < 0) data_bytes = 0;
        const long long pcm_frames = data_bytes / (2L * C);
        const long long nframes = pcm_frames / bi_.num_in + 1;                 /* the reference's rule, per channel */
        const long long n_in = nframes * bi_.num_in, cap = nframes * bi_.num_out;
        short *xi = calloc((size_t)n_in * C, 2), *xo = malloc((size_t)cap * C * 2);
        if (!xi || !xo) { fprintf(stderr, "out of memory\n"); return -1; }
        if (fread(xi, 2 * (size_t)C, (size_t)pcm_frames, fi) != (size_t)pcm_frames) { fprintf(stderr, "short read\n"); return -1; }
        long long n_out = 0;
        if (llz_cuda_resample_bank_run_pcm_host(hb, xi, LLZ_CUDA_PCM_S16, n_in, xo, LLZ_CUDA_PCM_S16, cap, &n_out) != 0) {
            fprintf(stderr, "run failed: %s\n", llz_cuda_last_error());
            return -1;
        }
        w.samplerate = type == 0 ? w.samplerate / down : type == 1 ? w.samplerate * up : (uint32_t)(((uint64_t)w.samplerate * up) / down);
        wav_write_header(fo, &w, (uint32_t)(n_out * C * 2));
        fseek(fo, 44, SEEK_SET);
        fwrite(xo, 2 * (size_t)C, (size_t)n_out, fo);
        if (!quiet) printf("frames = %lld, channels = %d, output bytes = %lld\n", nframes, C, n_out * C * 2);
        llz_resample_filter_uninit(hb);
        free(xi); free(xo);
        fclose(fi); fclose(fo);
        return 0;
    }

    unsigned long h;
    uint32_t rate_out;
    switch (type) {
    case 0:  h = llz_decimate_init(down, gain, BLACKMAN);             rate_out = w.samplerate / down; break;
    case 1:  h = llz_interp_init(up, gain, BLACKMAN);                 rate_out = w.samplerate * up; break;
    default: h = llz_resample_filter_init(up, down, gain, BLACKMAN);  rate_out = (uint32_t)(((uint64_t)w.samplerate * up) / down); break;
    }
    if (h == (unsigned long)-1) { fprintf(stderr, "init failed: %s\n", llz_cuda_last_error()); return -1; }

    llz_cuda_resample_info_t info;
    llz_cuda_resample_bank_info(h, &info);
    int in_bytes = llz_get_resample_framelen_bytes(h), out_bytes = 0;
    unsigned char *bi = malloc((size_t)in_bytes), *bo = malloc((size_t)info.num_out * 2);
    w.samplerate = rate_out;
    wav_write_header(fo, &w, 0);

    uint32_t total = 0;
    int frames = 0;
    if (whole) {
        /* same frame count as the loop below: floor(bytes / in_bytes) + 1, zero padded (main.c:91-119) */
        long pos = ftell(fi);
        fseek(fi, 0, SEEK_END);
        long data_bytes = ftell(fi) - pos;
        fseek(fi, pos, SEEK_SET);
        if (data_bytes < 0) data_bytes = 0;
        frames = (int)(data_bytes / in_bytes) + 1;
        size_t n_in = (size_t)frames * info.num_in, n_out_cap = (size_t)frames * info.num_out;
        short *xi = calloc(n_in, 2), *xo = malloc(n_out_cap * 2);
        if (!xi || !xo) { fprintf(stderr, "out of memory\n"); return -1; }
        if (fread(xi, 1, (size_t)data_bytes, fi) != (size_t)data_bytes) { fprintf(stderr, "short read\n"); return -1; }
        long long n_out = 0;
        if (llz_cuda_resample_bank_run_host(h, xi, (long long)n_in, (long long)n_in, xo, (long long)n_out_cap, &n_out) != 0) {
            fprintf(stderr, "run failed: %s\n", llz_cuda_last_error());
            return -1;
        }
        fwrite(xo, 2, (size_t)n_out, fo);
        total = (uint32_t)(n_out * 2);
        free(xi); free(xo);
    }
    for (int last = whole; !last;) {
        memset(bi, 0, (size_t)in_bytes);
        if ((int)fread(bi, 1, (size_t)in_bytes, fi) < in_bytes) last = 1;
        int rc = type == 0 ? llz_decimate(h, bi, in_bytes, bo, &out_bytes)
               : type == 1 ? llz_interp(h, bi, in_bytes, bo, &out_bytes)
                           : llz_resample(h, bi, in_bytes, bo, &out_bytes);
        if (rc != 0) { fprintf(stderr, "frame %d failed: %s\n", frames, llz_cuda_last_error()); return -1; }
        fwrite(bo, 1, (size_t)out_bytes, fo);
        total += (uint32_t)out_bytes;
        frames++;
    }
    if (!quiet) printf("frames = %d, output bytes = %u\n", frames, total);
    /* the reference stores data_size in blocks and writes blocks*block_align (main.c:121, llz_wavfmt.c:204,213) */
    uint32_t align = (uint32_t)w.channels * w.bytes_per_sample;
    wav_write_header(fo, &w, align ? total / align * align : total);
    llz_resample_filter_uninit(h);
    free(bi); free(bo);
    fclose(fi); fclose(fo);
    return 0;
}

/*
 * vit_oracle.c — CPU fp32 restatement of the ViT.rs hot path.  TEST INFRASTRUCTURE ONLY
 * (see vit_oracle.h).  "parity unpinned" beyond the reference's two exact known answers;
 * pinned here by finite differences and by PyTorch CPU fp32 golden vectors.
 *
 * Written from the reference's behaviour, op by op.  The op decomposition, signatures,
 * memory layouts, constants and accumulation order of the reference are kept so that fp32
 * results are reproducible; rows / heads that own disjoint accumulators may run on separate
 * OpenMP threads, which does not change any summation order.
 *
 * Citations are to /root/reference (train_vit.rs = tv, rusty_vit.rs = rv, attention.rs = at).
 */
#include "vit_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define GELU_K 0.044715f

/* ------------------------------------------------------------------------------------ */
/* residual: tv:376-382 / rv:460-470 (forward), tv:521-528 / rv:670-677 (backward)      */
void residual_forward(float* out, const float* inp1, const float* inp2, int n) {
#pragma omp parallel for if (n > 65536)
    for (int i = 0; i < n; ++i) out[i] = inp1[i] + inp2[i];
}

void residual_backward(float* dinp1, float* dinp2, const float* dout, int n) {
#pragma omp parallel for if (n > 65536)
    for (int i = 0; i < n; ++i) {
        float g = dout[i];
        dinp1[i] += g;
        dinp2[i] += g;
    }
}

/* ------------------------------------------------------------------------------------ */
/* matmul: out[bt,o] = bias[o] + sum_i inp[bt,i] * weight[o,i]; weight is [oc, c].
 * tv:384-398 / rv:484-498.  The accumulator starts at the bias and sums i ascending.   */
void matmul_forward(float* out, const float* inp, const float* weight, const float* bias,
                    int b, int t, int c, int oc) {
    const long rows = (long)b * t;
#pragma omp parallel for
    for (long r = 0; r < rows; ++r) {
        const float* x = inp + r * c;
        float* y = out + r * oc;
        for (int o = 0; o < oc; ++o) {
            const float* w = weight + (long)o * c;
            float acc = bias ? bias[o] : 0.0f;
            for (int i = 0; i < c; ++i) acc += x[i] * w[i];
            y[o] = acc;
        }
    }
}

/* tv:530-557 / rv:693-720.  Pass 1: dinp[bt,:] += sum_o weight[o,:]*dout[bt,o] (o ascending
 * per row).  Pass 2: for each o, over bt ascending: dbias[o] += dout, dweight[o,:] += inp*dout.
 * All three outputs accumulate.  dbias may be NULL (tv:548).                              */
void matmul_backward(float* dinp, float* dweight, float* dbias, const float* dout,
                     const float* inp, const float* weight, int b, int t, int c, int oc) {
    const long rows = (long)b * t;
#pragma omp parallel for
    for (long r = 0; r < rows; ++r) {
        float* dx = dinp + r * c;
        const float* dy = dout + r * oc;
        for (int o = 0; o < oc; ++o) {
            const float* w = weight + (long)o * c;
            const float g = dy[o];
            for (int i = 0; i < c; ++i) dx[i] += w[i] * g;
        }
    }
#pragma omp parallel for
    for (int o = 0; o < oc; ++o) {
        float* dw = dweight + (long)o * c;
        for (long r = 0; r < rows; ++r) {
            const float g = dout[r * oc + o];
            const float* x = inp + r * c;
            if (dbias) dbias[o] += g;
            for (int i = 0; i < c; ++i) dw[i] += x[i] * g;
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* attention: tv:400-451 / rv:512-563 / at:1-58.  inp is packed qkv [B,T,3C]: Q at column
 * h*hs, K at C + h*hs, V at 2C + h*hs.  scale = 1/sqrt(hs).  Scores are max-subtracted,
 * exponentiated, normalised by 1/sum (0 when the sum is 0, tv:432), then out = att . V.
 * Deviations: rows are (b*T + t) and the score buffers are [B,NH,T,T] (D2: the reference's
 * shadowed loop variables, tv:407-410); every weight incl. the diagonal is normalised
 * (D3: tv:434 stops one short); the running max starts at -inf as in at:22 (D6).
 * causal=1 keeps the reference's t2 <= t range; masked entries are stored as 0.          */
void attention_forward_ex(float* out, float* preatt, float* att, const float* inp,
                          int b, int t, int c, int nh, int causal) {
    const int c3 = 3 * c;
    const int hs = c / nh;
    const float scale = 1.0f / sqrtf((float)hs);
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi) {
        for (int h = 0; h < nh; ++h) {
            for (int tq = 0; tq < t; ++tq) {
                const float* q = inp + ((long)bi * t + tq) * c3 + h * hs;
                float* s_row = preatt + (((long)bi * nh + h) * t + tq) * t;
                float* p_row = att + (((long)bi * nh + h) * t + tq) * t;
                const int kend = causal ? tq + 1 : t;

                float mx = -INFINITY;
                for (int tk = 0; tk < kend; ++tk) {
                    const float* k = inp + ((long)bi * t + tk) * c3 + c + h * hs;
                    float dot = 0.0f;
                    for (int i = 0; i < hs; ++i) dot += q[i] * k[i];
                    dot *= scale;
                    if (dot > mx) mx = dot;
                    s_row[tk] = dot;
                }
                float sum = 0.0f;
                for (int tk = 0; tk < kend; ++tk) {
                    float e = expf(s_row[tk] - mx);
                    sum += e;
                    p_row[tk] = e;
                }
                const float inv = (sum == 0.0f) ? 0.0f : 1.0f / sum;
                for (int tk = 0; tk < kend; ++tk) p_row[tk] *= inv;
                for (int tk = kend; tk < t; ++tk) { s_row[tk] = 0.0f; p_row[tk] = 0.0f; }

                float* o = out + ((long)bi * t + tq) * c + h * hs;
                for (int i = 0; i < hs; ++i) o[i] = 0.0f;
                for (int tk = 0; tk < kend; ++tk) {
                    const float* v = inp + ((long)bi * t + tk) * c3 + 2 * c + h * hs;
                    const float p = p_row[tk];
                    for (int i = 0; i < hs; ++i) o[i] += p * v[i];
                }
            }
        }
    }
}

void attention_forward(float* out, float* preatt, float* att, const float* inp,
                       int b, int t, int c, int nh) {
    attention_forward_ex(out, preatt, att, inp, b, t, c, nh, 1);
}

/* tv:559-601 (the only definition).  Per query row: datt += V.dout, dV += att*dout
 * (tv:574-581); dpreatt[t3] += att[t2]*(delta(t2,t3) - att[t3])*datt[t2] (tv:583-589);
 * dQ += K*dpreatt*scale, dK += Q*dpreatt*scale (tv:591-598).  Everything accumulates.    */
void attention_backward_ex(float* dinp, float* dpreatt, float* datt, const float* dout,
                           const float* inp, const float* att, int b, int t, int c, int nh,
                           int causal) {
    const int c3 = 3 * c;
    const int hs = c / nh;
    const float scale = 1.0f / sqrtf((float)hs);
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi) {
        for (int h = 0; h < nh; ++h) {
            for (int tq = 0; tq < t; ++tq) {
                const long row = ((long)bi * nh + h) * t + tq;
                const float* p_row = att + row * t;
                float* dp_row = datt + row * t;
                float* ds_row = dpreatt + row * t;
                const float* q = inp + ((long)bi * t + tq) * c3 + h * hs;
                float* dq = dinp + ((long)bi * t + tq) * c3 + h * hs;
                const float* dy = dout + ((long)bi * t + tq) * c + h * hs;
                const int kend = causal ? tq + 1 : t;

                for (int tk = 0; tk < kend; ++tk) {
                    const float* v = inp + ((long)bi * t + tk) * c3 + 2 * c + h * hs;
                    float* dv = dinp + ((long)bi * t + tk) * c3 + 2 * c + h * hs;
                    for (int i = 0; i < hs; ++i) {
                        dp_row[tk] += v[i] * dy[i];
                        dv[i] += p_row[tk] * dy[i];
                    }
                }
                for (int t2 = 0; t2 < kend; ++t2) {
                    for (int t3 = 0; t3 < kend; ++t3) {
                        const float ind = (t2 == t3) ? 1.0f : 0.0f;
                        const float local = p_row[t2] * (ind - p_row[t3]);
                        ds_row[t3] += local * dp_row[t2];
                    }
                }
                for (int tk = 0; tk < kend; ++tk) {
                    const float* k = inp + ((long)bi * t + tk) * c3 + c + h * hs;
                    float* dk = dinp + ((long)bi * t + tk) * c3 + c + h * hs;
                    const float g = ds_row[tk] * scale;
                    for (int i = 0; i < hs; ++i) {
                        dq[i] += k[i] * g;
                        dk[i] += q[i] * g;
                    }
                }
            }
        }
    }
}

void attention_backward(float* dinp, float* dpreatt, float* datt, const float* dout,
                        const float* inp, const float* att, int b, int t, int c, int nh) {
    attention_backward_ex(dinp, dpreatt, datt, dout, inp, att, b, t, c, nh, 1);
}

/* ------------------------------------------------------------------------------------ */
/* layernorm: tv:453-480 / rv:578-605.  Two-pass mean and biased variance, eps = 1e-5,
 * rstd = 1/sqrt(var+eps), out = (x-mean)*rstd*w + b; mean and rstd are saved.           */
void layernorm_forward(float* out, float* mean, float* rstd, const float* inp,
                       const float* weight, const float* bias, int b, int t, int c) {
    const float eps = 1e-5f;
    const long rows = (long)b * t;
#pragma omp parallel for
    for (long r = 0; r < rows; ++r) {
        const float* x = inp + r * c;
        float m = 0.0f;
        for (int i = 0; i < c; ++i) m += x[i];
        m /= (float)c;
        float var = 0.0f;
        for (int i = 0; i < c; ++i) {
            float d = x[i] - m;
            var += d * d;
        }
        var /= (float)c;
        const float s = 1.0f / sqrtf(var + eps);
        float* y = out + r * c;
        for (int i = 0; i < c; ++i) y[i] = (s * (x[i] - m)) * weight[i] + bias[i];
        mean[r] = m;
        rstd[r] = s;
    }
}

/* tv:603-637 / rv:737-783.  dnorm = w*dout; dx += (dnorm - mean(dnorm)
 * - norm*mean(dnorm*norm))*rstd; dw += norm*dout; db += dout.  Rows ascending, so dw/db
 * sum in row order.  (tv:615 lacks a dereference; the evident read of inp[i] is used.)   */
void layernorm_backward(float* dinp, float* dweight, float* dbias, const float* dout,
                        const float* inp, const float* weight, const float* mean,
                        const float* rstd, int b, int t, int c) {
    const long rows = (long)b * t;
    for (long r = 0; r < rows; ++r) {
        const float* dy = dout + r * c;
        const float* x = inp + r * c;
        float* dx = dinp + r * c;
        const float m = mean[r];
        const float s = rstd[r];
        float dn_mean = 0.0f, dnn_mean = 0.0f;
        for (int i = 0; i < c; ++i) {
            float nrm = (x[i] - m) * s;
            float dn = weight[i] * dy[i];
            dn_mean += dn;
            dnn_mean += dn * nrm;
        }
        dn_mean /= (float)c;
        dnn_mean /= (float)c;
        for (int i = 0; i < c; ++i) {
            float nrm = (x[i] - m) * s;
            float dn = weight[i] * dy[i];
            dbias[i] += dy[i];
            dweight[i] += nrm * dy[i];
            float dv = 0.0f;
            dv += dn;
            dv -= dn_mean;
            dv -= nrm * dnn_mean;
            dv *= s;
            dx[i] += dv;
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* gelu (tanh form): tv:482-491 / rv:614-623.                                            */
void gelu_forward(float* out, const float* inp, int n) {
    const float s = sqrtf(2.0f / (float)M_PI);
#pragma omp parallel for if (n > 65536)
    for (int i = 0; i < n; ++i) {
        float x = inp[i];
        float cube = GELU_K * x * x * x;
        out[i] = 0.5f * x * (1.0f + tanhf(s * (x + cube)));
    }
}

/* tv:639-653 / rv:793-807 with the derivative corrected (D4): sech^2(u) = 1/cosh(u)^2,
 * where the reference evaluates cosh at 2u and is then not the derivative of its forward. */
void gelu_backward(float* dinp, const float* inp, const float* dout, int n) {
    const float s = sqrtf(2.0f / (float)M_PI);
#pragma omp parallel for if (n > 65536)
    for (int i = 0; i < n; ++i) {
        float x = inp[i];
        float cube = GELU_K * x * x * x;
        float u = s * (x + cube);
        float th = tanhf(u);
        float ch = coshf(u);
        float sech2 = 1.0f / (ch * ch);
        float local = 0.5f * (1.0f + th) + x * 0.5f * sech2 * s * (1.0f + 3.0f * GELU_K * x * x);
        dinp[i] += local * dout[i];
    }
}

/* ------------------------------------------------------------------------------------ */
/* softmax over the last axis: tv:493-517 / rv:634-658 (max start -inf, D6).             */
void softmax_forward(float* probs, const float* logits, int b, int t, int v) {
    const long rows = (long)b * t;
#pragma omp parallel for if (rows * v > 65536)
    for (long r = 0; r < rows; ++r) {
        const float* z = logits + r * v;
        float* p = probs + r * v;
        float mx = -INFINITY;
        for (int i = 0; i < v; ++i)
            if (z[i] > mx) mx = z[i];
        float sum = 0.0f;
        for (int i = 0; i < v; ++i) {
            p[i] = expf(z[i] - mx);
            sum += p[i];
        }
        for (int i = 0; i < v; ++i) p[i] /= sum;
    }
}

/* rv:836-843 with the evident loss (D5): losses[i] = -ln probs[i, target_i].            */
void crossentropy_forward(float* losses, const float* probs, const int* targets,
                          int b, int t, int v) {
    const long rows = (long)b * t;
    for (long r = 0; r < rows; ++r) losses[r] = -logf(probs[r * v + targets[r]]);
}

/* Called at rv:371 / tv:293, never defined in the reference; by signature (D5):
 * dlogits[i,j] += (probs[i,j] - [j == target_i]) * dlosses[i].                          */
void crossentropy_softmax_backward(float* dlogits, const float* dlosses, const float* probs,
                                   const int* targets, int b, int t, int v) {
    const long rows = (long)b * t;
    for (long r = 0; r < rows; ++r) {
        const float g = dlosses[r];
        const int tgt = targets[r];
        for (int j = 0; j < v; ++j) {
            float ind = (j == tgt) ? 1.0f : 0.0f;
            dlogits[r * v + j] += (probs[r * v + j] - ind) * g;
        }
    }
}

/* ------------------------------------------------------------------------------------ */
/* token + position embedding, by the signature used at rv:282 / rv:448.                 */
void encoder_forward(float* encoded, const int* inputs, const float* wte, const float* wpe,
                     int b, int t, int c) {
    for (int bi = 0; bi < b; ++bi)
        for (int ti = 0; ti < t; ++ti) {
            float* o = encoded + ((long)bi * t + ti) * c;
            const float* e = wte + (long)inputs[bi * t + ti] * c;
            const float* p = wpe + (long)ti * c;
            for (int i = 0; i < c; ++i) o[i] = e[i] + p[i];
        }
}

void encoder_backward(float* dwte, float* dwpe, const float* dencoded, const int* inputs,
                      int b, int t, int c) {
    for (int bi = 0; bi < b; ++bi)
        for (int ti = 0; ti < t; ++ti) {
            const float* g = dencoded + ((long)bi * t + ti) * c;
            float* de = dwte + (long)inputs[bi * t + ti] * c;
            float* dp = dwpe + (long)ti * c;
            for (int i = 0; i < c; ++i) {
                de[i] += g[i];
                dp[i] += g[i];
            }
        }
}

/* ------------------------------------------------------------------------------------ */
/* patch embedding (D7: replaces encoder_forward for a ViT; written in matmul_forward's
 * accumulation order: accumulator starts at the bias, k = (ch, i, j) ascending, then the
 * position row is added).                                                                */
void patch_embed_forward(float* encoded, const float* images, const float* patchw,
                         const float* patchb, const float* cls, const float* wpe,
                         int b, int img, int patch, int c) {
    const int g = img / patch;       /* patches per side */
    const int np = g * g;
    const int t = np + 1;
    const int kdim = 3 * patch * patch;
#pragma omp parallel for
    for (int bi = 0; bi < b; ++bi) {
        float* row0 = encoded + (long)bi * t * c;
        for (int o = 0; o < c; ++o) row0[o] = cls[o] + wpe[o];
        for (int n = 0; n < np; ++n) {
            const int py = n / g, px = n % g;
            float* y = encoded + ((long)bi * t + 1 + n) * c;
            const float* pos = wpe + (long)(1 + n) * c;
            for (int o = 0; o < c; ++o) {
                const float* w = patchw + (long)o * kdim;
                float acc = patchb[o];
                for (int ch = 0; ch < 3; ++ch)
                    for (int i = 0; i < patch; ++i) {
                        const float* px_row = images + (((long)bi * 3 + ch) * img + (py * patch + i)) * img + px * patch;
                        const float* wk = w + (ch * patch + i) * patch;
                        for (int j = 0; j < patch; ++j) acc += px_row[j] * wk[j];
                    }
                y[o] = acc + pos[o];
            }
        }
    }
}

/* matmul_backward's second pass applied to the patch rows (no dinp: images are data),
 * plus the position/CLS scatter of encoder_backward.  All outputs accumulate.            */
void patch_embed_backward(float* dpatchw, float* dpatchb, float* dcls, float* dwpe,
                          const float* dencoded, const float* images,
                          int b, int img, int patch, int c) {
    const int g = img / patch;
    const int np = g * g;
    const int t = np + 1;
    const int kdim = 3 * patch * patch;
    for (int bi = 0; bi < b; ++bi)
        for (int ti = 0; ti < t; ++ti) {
            const float* gr = dencoded + ((long)bi * t + ti) * c;
            float* dp = dwpe + (long)ti * c;
            for (int o = 0; o < c; ++o) dp[o] += gr[o];
            if (ti == 0)
                for (int o = 0; o < c; ++o) dcls[o] += gr[o];
        }
#pragma omp parallel for
    for (int o = 0; o < c; ++o) {
        float* dw = dpatchw + (long)o * kdim;
        for (int bi = 0; bi < b; ++bi)
            for (int n = 0; n < np; ++n) {
                const int py = n / g, px = n % g;
                const float gv = dencoded[((long)bi * t + 1 + n) * c + o];
                dpatchb[o] += gv;
                for (int ch = 0; ch < 3; ++ch)
                    for (int i = 0; i < patch; ++i) {
                        const float* px_row = images + (((long)bi * 3 + ch) * img + (py * patch + i)) * img + px * patch;
                        float* dwk = dw + (ch * patch + i) * patch;
                        for (int j = 0; j < patch; ++j) dwk[j] += px_row[j] * gv;
                    }
            }
    }
}

void cls_gather(float* out, const float* inp, int b, int t, int c) {
    for (int bi = 0; bi < b; ++bi) memcpy(out + (long)bi * c, inp + (long)bi * t * c, sizeof(float) * c);
}

void cls_scatter_add(float* dinp, const float* dout, int b, int t, int c) {
    for (int bi = 0; bi < b; ++bi)
        for (int i = 0; i < c; ++i) dinp[(long)bi * t * c + i] += dout[(long)bi * c + i];
}

/* ------------------------------------------------------------------------------------ */
/* optimiser.  The reference ships SGD (tv:737-743) and allocates unused m/v (rv:67-68);
 * AdamW is the north-star's optimiser (D8), in torch.optim.AdamW's operation order.     */
void sgd_step(float* params, const float* grads, size_t n, float lr) {
    for (size_t i = 0; i < n; ++i) params[i] -= lr * grads[i];
}

void adamw_step(float* params, const float* grads, float* m, float* v, size_t n,
                float lr, float beta1, float beta2, float eps, float weight_decay, int step) {
    const float bc1 = 1.0f - powf(beta1, (float)step);
    const float bc2 = 1.0f - powf(beta2, (float)step);
    const float step_size = lr / bc1;
    const float bc2_sqrt = sqrtf(bc2);
#pragma omp parallel for if (n > 65536)
    for (size_t i = 0; i < n; ++i) {
        float g = grads[i];
        float p = params[i];
        p = p * (1.0f - lr * weight_decay);
        float mi = beta1 * m[i] + (1.0f - beta1) * g;
        float vi = beta2 * v[i] + (1.0f - beta2) * g * g;
        m[i] = mi;
        v[i] = vi;
        float denom = sqrtf(vi) / bc2_sqrt + eps;
        params[i] = p - step_size * (mi / denom);
    }
}

/* ------------------------------------------------------------------------------------ */
/* counter-based uniform generator (D9): splitmix64 finaliser over (seed, stream, idx).   */
float vit_rand_u01(uint64_t seed, uint64_t stream, uint64_t idx) {
    uint64_t x = seed * 0x9E3779B97F4A7C15ull + stream * 0xD1B54A32D192ED03ull + idx;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (float)(x >> 40) * (1.0f / 16777216.0f);
}

void vit_fill_uniform(float* dst, size_t n, uint64_t seed, uint64_t stream, float lo, float hi) {
    for (size_t i = 0; i < n; ++i) dst[i] = lo + (hi - lo) * vit_rand_u01(seed, stream, i);
}

/* ------------------------------------------------------------------------------------ */
/* model: rv:63-259 (storage), rv:269-351 (forward), rv:354-449 (backward).              */
void vit_param_sizes(const ViTConfig* cfg, size_t* s) {
    const size_t c = cfg->channels, l = cfg->num_layers, t = cfg->max_seq_len;
    const size_t kdim = 3u * cfg->patch_size * cfg->patch_size;
    s[0] = c * kdim;            /* patchw  (replaces wte, rv:105) */
    s[1] = c;                   /* patchb */
    s[2] = c;                   /* cls */
    s[3] = t * c;               /* wpe     (rv:106) */
    s[4] = l * c;               /* ln1w    (rv:107-121 follow) */
    s[5] = l * c;               /* ln1b */
    s[6] = l * 3 * c * c;       /* qkvw */
    s[7] = l * 3 * c;           /* qkvb */
    s[8] = l * c * c;           /* attprojw */
    s[9] = l * c;               /* attprojb */
    s[10] = l * c;              /* ln2w */
    s[11] = l * c;              /* ln2b */
    s[12] = l * 4 * c * c;      /* fcw */
    s[13] = l * 4 * c;          /* fcb */
    s[14] = l * c * 4 * c;      /* fcprojw */
    s[15] = l * c;              /* fcprojb */
    s[16] = c;                  /* lnfw */
    s[17] = c;                  /* lnfb */
    s[18] = (size_t)cfg->num_classes * c; /* headw (untied head, D7) */
    s[19] = (size_t)cfg->num_classes;     /* headb */
}

static void carve(float* base, const size_t* sizes, int n, float** views) {
    for (int i = 0; i < n; ++i) {
        views[i] = base;
        base += sizes[i];
    }
}

ViT* vit_build(const ViTConfig* cfg, uint64_t seed, int init_mode) {
    ViT* m = (ViT*)calloc(1, sizeof(ViT));
    m->config = *cfg;
    vit_param_sizes(cfg, m->param_sizes);
    m->num_parameters = 0;
    for (int i = 0; i < VIT_NUM_PARAMETER_TENSORS; ++i) m->num_parameters += m->param_sizes[i];
    m->params_memory = (float*)calloc(m->num_parameters, sizeof(float));
    m->grads_memory = (float*)calloc(m->num_parameters, sizeof(float));
    m->m_memory = (float*)calloc(m->num_parameters, sizeof(float));
    m->v_memory = (float*)calloc(m->num_parameters, sizeof(float));
    carve(m->params_memory, m->param_sizes, VIT_NUM_PARAMETER_TENSORS, (float**)&m->params);
    carve(m->grads_memory, m->param_sizes, VIT_NUM_PARAMETER_TENSORS, (float**)&m->grads);
    m->mean_loss = -1.0f;

    /* init: rv:864-903 — weights U[0,1)*0.02, LN gains 1, every bias 0 (D9).
     * init_mode 1 = symmetric U[-1,1)*0.02 weights (build-defined alternative). */
    const float lo = init_mode == 1 ? -0.02f : 0.0f, hi = 0.02f;
    float** p = (float**)&m->params;
    const int weight_ids[] = {0, 2, 3, 6, 8, 12, 14, 18};
    for (unsigned k = 0; k < sizeof(weight_ids) / sizeof(int); ++k) {
        int id = weight_ids[k];
        vit_fill_uniform(p[id], m->param_sizes[id], seed, (uint64_t)id, lo, hi);
    }
    const int gain_ids[] = {4, 10, 16};
    for (unsigned k = 0; k < 3; ++k)
        for (size_t i = 0; i < m->param_sizes[gain_ids[k]]; ++i) p[gain_ids[k]][i] = 1.0f;
    return m;
}

static void alloc_acts(ViT* m, int b) {
    const size_t B = b, T = m->config.max_seq_len, C = m->config.channels, L = m->config.num_layers,
                 NH = m->config.num_heads, V = m->config.num_classes;
    size_t* s = m->act_sizes;
    /* rv:150-174 with the batch factor restored */
    s[0] = B * T * C;             /* encoded */
    s[1] = L * B * T * C;         /* ln1 */
    s[2] = L * B * T;             /* ln1_mean */
    s[3] = L * B * T;             /* ln1_rstd */
    s[4] = L * B * T * 3 * C;     /* qkv */
    s[5] = L * B * T * C;         /* atty */
    s[6] = L * B * NH * T * T;    /* preatt */
    s[7] = L * B * NH * T * T;    /* att */
    s[8] = L * B * T * C;         /* attproj */
    s[9] = L * B * T * C;         /* residual2 */
    s[10] = L * B * T * C;        /* ln2 */
    s[11] = L * B * T;            /* ln2_mean */
    s[12] = L * B * T;            /* ln2_rstd */
    s[13] = L * B * T * 4 * C;    /* fch */
    s[14] = L * B * T * 4 * C;    /* fch_gelu */
    s[15] = L * B * T * C;        /* fcproj */
    s[16] = L * B * T * C;        /* residual3 */
    s[17] = B * C;                /* lnf      (CLS rows only, D7) */
    s[18] = B;                    /* lnf_mean */
    s[19] = B;                    /* lnf_rstd */
    s[20] = B * V;                /* logits */
    s[21] = B * V;                /* probs */
    s[22] = B;                    /* losses */
    m->num_activations = 0;
    for (int i = 0; i < VIT_NUM_ACTIVATION_TENSORS; ++i) m->num_activations += s[i];
    free(m->acts_memory);
    free(m->grads_acts_memory);
    free(m->targets);
    /* + B*C scratch at the end of each arena: the gathered CLS rows and their gradient */
    m->acts_memory = (float*)calloc(m->num_activations + B * C, sizeof(float));
    m->grads_acts_memory = (float*)calloc(m->num_activations + B * C, sizeof(float));
    carve(m->acts_memory, s, VIT_NUM_ACTIVATION_TENSORS, (float**)&m->acts);
    carve(m->grads_acts_memory, s, VIT_NUM_ACTIVATION_TENSORS, (float**)&m->grads_acts);
    m->targets = (int*)calloc(B, sizeof(int));
    m->batch_size = b;
    m->seq_len = (int)T;
}

void vit_forward(ViT* m, const float* images, const int* targets, int b) {
    if (m->acts_memory == NULL || b != m->batch_size) alloc_acts(m, b);
    const int T = m->config.max_seq_len, C = m->config.channels, L = m->config.num_layers,
              NH = m->config.num_heads, V = m->config.num_classes;
    const ParameterTensors* P = &m->params;
    ActivationTensors* A = &m->acts;
    const long btc = (long)b * T * C;
    m->inputs = images;
    if (targets) memcpy(m->targets, targets, sizeof(int) * b);

    patch_embed_forward(A->encoded, images, P->patchw, P->patchb, P->cls, P->wpe,
                        b, m->config.image_size, m->config.patch_size, C);
    const float* residual = A->encoded;
    for (int l = 0; l < L; ++l) {
        residual = l == 0 ? A->encoded : A->residual3 + (l - 1) * btc;
        float* ln1 = A->ln1 + l * btc;
        float* qkv = A->qkv + l * btc * 3;
        float* atty = A->atty + l * btc;
        float* preatt = A->preatt + (long)l * b * NH * T * T;
        float* att = A->att + (long)l * b * NH * T * T;
        float* attproj = A->attproj + l * btc;
        float* residual2 = A->residual2 + l * btc;
        float* ln2 = A->ln2 + l * btc;
        float* fch = A->fch + l * btc * 4;
        float* fch_gelu = A->fch_gelu + l * btc * 4;
        float* fcproj = A->fcproj + l * btc;
        float* residual3 = A->residual3 + l * btc;
        /* op order: rv:322-331 */
        layernorm_forward(ln1, A->ln1_mean + (long)l * b * T, A->ln1_rstd + (long)l * b * T, residual,
                          P->ln1w + l * C, P->ln1b + l * C, b, T, C);
        matmul_forward(qkv, ln1, P->qkvw + (long)l * 3 * C * C, P->qkvb + l * 3 * C, b, T, C, 3 * C);
        attention_forward_ex(atty, preatt, att, qkv, b, T, C, NH, m->config.causal);
        matmul_forward(attproj, atty, P->attprojw + (long)l * C * C, P->attprojb + l * C, b, T, C, C);
        residual_forward(residual2, residual, attproj, (int)btc);
        layernorm_forward(ln2, A->ln2_mean + (long)l * b * T, A->ln2_rstd + (long)l * b * T, residual2,
                          P->ln2w + l * C, P->ln2b + l * C, b, T, C);
        matmul_forward(fch, ln2, P->fcw + (long)l * 4 * C * C, P->fcb + l * 4 * C, b, T, C, 4 * C);
        gelu_forward(fch_gelu, fch, (int)(btc * 4));
        matmul_forward(fcproj, fch_gelu, P->fcprojw + (long)l * C * 4 * C, P->fcprojb + l * C, b, T, 4 * C, C);
        residual_forward(residual3, residual2, fcproj, (int)btc);
    }
    /* head (rv:335-347, on the CLS row of each image, D7) */
    float* cls_rows = m->acts_memory + m->num_activations;
    cls_gather(cls_rows, A->residual3 + (L - 1) * btc, b, T, C);
    layernorm_forward(A->lnf, A->lnf_mean, A->lnf_rstd, cls_rows, P->lnfw, P->lnfb, b, 1, C);
    matmul_forward(A->logits, A->lnf, P->headw, P->headb, b, 1, C, V);
    softmax_forward(A->probs, A->logits, b, 1, V);
    if (targets) {
        crossentropy_forward(A->losses, A->probs, m->targets, b, 1, V);
        float mean_loss = 0.0f;
        for (int i = 0; i < b; ++i) mean_loss += A->losses[i];
        m->mean_loss = mean_loss / (float)b;
    } else {
        m->mean_loss = -1.0f;   /* rv:348-350 */
    }
}

void vit_zero_grad(ViT* m) {
    memset(m->grads_memory, 0, sizeof(float) * m->num_parameters);
    if (m->grads_acts_memory)
        memset(m->grads_acts_memory, 0, sizeof(float) * (m->num_activations + (size_t)m->batch_size * m->config.channels));
}

void vit_backward(ViT* m) {
    const int b = m->batch_size, T = m->config.max_seq_len, C = m->config.channels,
              L = m->config.num_layers, NH = m->config.num_heads, V = m->config.num_classes;
    const ParameterTensors* P = &m->params;
    ParameterTensors* G = &m->grads;
    const ActivationTensors* A = &m->acts;
    ActivationTensors* D = &m->grads_acts;
    const long btc = (long)b * T * C;

    /* rv:366-369: every sample's loss gets 1/(number of loss rows); a data-parallel caller
     * passes 1/B_global so a sum all-reduce reproduces the single-process gradient. */
    const float dloss = m->dloss_scale != 0.0f ? m->dloss_scale : 1.0f / (float)b;
    for (int i = 0; i < b; ++i) D->losses[i] = dloss;

    crossentropy_softmax_backward(D->logits, D->losses, A->probs, m->targets, b, 1, V);
    matmul_backward(D->lnf, G->headw, G->headb, D->logits, A->lnf, P->headw, b, 1, C, V);
    const float* cls_rows = m->acts_memory + m->num_activations;
    float* dcls_rows = m->grads_acts_memory + m->num_activations;
    layernorm_backward(dcls_rows, G->lnfw, G->lnfb, D->lnf, cls_rows, P->lnfw, A->lnf_mean, A->lnf_rstd, b, 1, C);
    cls_scatter_add(D->residual3 + (L - 1) * btc, dcls_rows, b, T, C);

    for (int l = L - 1; l >= 0; --l) {
        const float* residual = l == 0 ? A->encoded : A->residual3 + (l - 1) * btc;
        float* dresidual = l == 0 ? D->encoded : D->residual3 + (l - 1) * btc;
        const long lbt = (long)l * b * T;
        const long latt = (long)l * b * NH * T * T;
        /* op order: rv:436-445 */
        residual_backward(D->residual2 + l * btc, D->fcproj + l * btc, D->residual3 + l * btc, (int)btc);
        matmul_backward(D->fch_gelu + l * btc * 4, G->fcprojw + (long)l * C * 4 * C, G->fcprojb + l * C,
                        D->fcproj + l * btc, A->fch_gelu + l * btc * 4, P->fcprojw + (long)l * C * 4 * C, b, T, 4 * C, C);
        gelu_backward(D->fch + l * btc * 4, A->fch + l * btc * 4, D->fch_gelu + l * btc * 4, (int)(btc * 4));
        matmul_backward(D->ln2 + l * btc, G->fcw + (long)l * 4 * C * C, G->fcb + l * 4 * C,
                        D->fch + l * btc * 4, A->ln2 + l * btc, P->fcw + (long)l * 4 * C * C, b, T, C, 4 * C);
        layernorm_backward(D->residual2 + l * btc, G->ln2w + l * C, G->ln2b + l * C, D->ln2 + l * btc,
                           A->residual2 + l * btc, P->ln2w + l * C, A->ln2_mean + lbt, A->ln2_rstd + lbt, b, T, C);
        residual_backward(dresidual, D->attproj + l * btc, D->residual2 + l * btc, (int)btc);
        matmul_backward(D->atty + l * btc, G->attprojw + (long)l * C * C, G->attprojb + l * C,
                        D->attproj + l * btc, A->atty + l * btc, P->attprojw + (long)l * C * C, b, T, C, C);
        attention_backward_ex(D->qkv + l * btc * 3, D->preatt + latt, D->att + latt, D->atty + l * btc,
                              A->qkv + l * btc * 3, A->att + latt, b, T, C, NH, m->config.causal);
        matmul_backward(D->ln1 + l * btc, G->qkvw + (long)l * 3 * C * C, G->qkvb + l * 3 * C,
                        D->qkv + l * btc * 3, A->ln1 + l * btc, P->qkvw + (long)l * 3 * C * C, b, T, C, 3 * C);
        layernorm_backward(dresidual, G->ln1w + l * C, G->ln1b + l * C, D->ln1 + l * btc, residual,
                           P->ln1w + l * C, A->ln1_mean + lbt, A->ln1_rstd + lbt, b, T, C);
    }
    patch_embed_backward(G->patchw, G->patchb, G->cls, G->wpe, D->encoded, m->inputs,
                         b, m->config.image_size, m->config.patch_size, C);
}

void vit_update(ViT* m, float lr, float beta1, float beta2, float eps, float weight_decay) {
    m->adam_step += 1;
    adamw_step(m->params_memory, m->grads_memory, m->m_memory, m->v_memory, m->num_parameters,
               lr, beta1, beta2, eps, weight_decay, m->adam_step);
}

void vit_free(ViT* m) {
    if (!m) return;
    free(m->params_memory); free(m->grads_memory); free(m->m_memory); free(m->v_memory);
    free(m->acts_memory); free(m->grads_acts_memory); free(m->targets);
    free(m);
}

float* vit_param_ptr(ViT* m, int idx) { return ((float**)&m->params)[idx]; }
float* vit_grad_ptr(ViT* m, int idx) { return ((float**)&m->grads)[idx]; }
float* vit_act_ptr(ViT* m, int idx) { return ((float**)&m->acts)[idx]; }
float* vit_grad_act_ptr(ViT* m, int idx) { return ((float**)&m->grads_acts)[idx]; }
size_t vit_act_size(ViT* m, int idx) { return m->act_sizes[idx]; }

float vit_mean_loss(const ViT* m) { return m->mean_loss; }
void vit_set_dloss_scale(ViT* m, float s) { m->dloss_scale = s; }

int vit_oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void vit_oracle_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

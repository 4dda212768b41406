// simsched.c -- scratch: warp-instruction cost of scheduling policies on recorded per-ray op traces.
// input (binary): int32 n, int32 max_seg, int32 nseg[n], uint16 seg[n*max_seg]; rays in work-item order
// (items 32k..32k+31 = one 8x4 tile for primary rays); nseg==0 -> item produces no ray (gated pixel).
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static double g_isteps, g_ilanes, g_lphases, g_llanes, g_ltris;
static int CI = 80, CT = 75, CGEN = 120, CLEAF = 12, CREFILL = 25, CSWITCH = 45, CSTORE = 8;

typedef struct { const uint16_t* seg; int nseg; int k; int rem_i; int rem_t; int state; } RayS;  // state 0 empty,1 inner,2 leaf
static void ray_load(RayS* r, const uint16_t* seg, int nseg) {
    r->seg = seg; r->nseg = nseg; r->k = 0;
    if (nseg == 0) { r->state = 0; return; }
    r->rem_i = seg[0]; r->rem_t = seg[1];
    r->state = r->rem_i > 0 ? 1 : 2;
}
static void ray_after_inner(RayS* r) {  // one inner visit done
    if (--r->rem_i > 0) return;
    if (r->rem_t > 0) { r->state = 2; return; }
    // trailing run with no leaf (tri count 0): ray ends, or leaf with 0 tris
    r->k += 2;
    if (r->k >= r->nseg) { r->state = 0; return; }
    r->rem_i = r->seg[r->k]; r->rem_t = r->seg[r->k + 1];
    r->state = r->rem_i > 0 ? 1 : 2;
}
static void ray_after_leaf(RayS* r) {  // whole leaf done
    r->k += 2;
    if (r->k >= r->nseg) { r->state = 0; return; }
    r->rem_i = r->seg[r->k]; r->rem_t = r->seg[r->k + 1];
    r->state = r->rem_i > 0 ? 1 : (r->rem_t > 0 ? 2 : 1);
    if (r->rem_i == 0 && r->rem_t == 0) ray_after_leaf(r);
}

int main(int argc, char** argv) {
    FILE* f = fopen(argv[1], "rb");
    int n, max_seg;
    fread(&n, 4, 1, f); fread(&max_seg, 4, 1, f);
    int* nseg = malloc(4 * (size_t)n);
    uint16_t* seg = malloc(2 * (size_t)n * max_seg);
    fread(nseg, 4, n, f); fread(seg, 2, (size_t)n * max_seg, f);
    fclose(f);
    double useful = 0;  // thread-instructions of pure traversal work
    long nrays = 0;
    for (int i = 0; i < n; i++) {
        if (!nseg[i]) continue;
        nrays++;
        for (int k = 0; k < nseg[i]; k += 2) useful += (double)seg[(size_t)i * max_seg + k] * CI + (double)seg[(size_t)i * max_seg + k + 1] * CT;
    }
    printf("items %d rays %ld useful thread-instr %.3e -> ideal warp-instr %.3e\n", n, nrays, useful, useful / 32);

    // ---- policy A: batch of 32, classic synced while-while (all lanes finish inner phase, then leaf phase)
    {
        double cost = 0;
        for (int b = 0; b + 32 <= n; b += 32) {
            RayS r[32]; int any = 0;
            for (int l = 0; l < 32; l++) { ray_load(&r[l], seg + (size_t)(b + l) * max_seg, nseg[b + l]); any |= r[l].state; }
            cost += CGEN;
            if (!any) continue;
            for (;;) {
                int ni = 0, nl = 0;
                for (int l = 0; l < 32; l++) { ni += r[l].state == 1; nl += r[l].state == 2; }
                if (!ni && !nl) break;
                while (ni) { cost += CI; g_isteps++; g_ilanes += ni; ni = 0; for (int l = 0; l < 32; l++) if (r[l].state == 1) { ray_after_inner(&r[l]); } for (int l = 0; l < 32; l++) ni += r[l].state == 1; }
                int maxt = 0; for (int l = 0; l < 32; l++) if (r[l].state == 2 && r[l].rem_t > maxt) maxt = r[l].rem_t;
                if (maxt) { cost += CLEAF + (double)maxt * CT; g_lphases++; g_ltris += maxt; for (int l = 0; l < 32; l++) g_llanes += r[l].state == 2; }
                for (int l = 0; l < 32; l++) if (r[l].state == 2) ray_after_leaf(&r[l]);
            }
            cost += CSTORE;
        }
        printf("A  batch synced while-while            : warp-instr %.3e  efficiency %.3f\n", cost, useful / 32 / cost);
        printf("   inner steps %.3e avg lanes %.1f | leaf phases %.3e avg lanes %.1f avg max tris %.2f\n", g_isteps, g_ilanes / g_isteps, g_lphases, g_llanes / g_lphases, g_ltris / g_lphases);
        g_isteps = g_ilanes = g_lphases = g_llanes = g_ltris = 0;
    }
    // ---- policy B: persistent lanes (refill threshold R, inner exit threshold E), 1 ray per lane
    int Rs[] = {16, 8}, Es[] = {8, 16};
    for (int ri = 0; ri < 2; ri++) for (int ei = 0; ei < 2; ei++) {
        int R = Rs[ri], E = Es[ei];
        double cost = 0; long next = 0;
        const int W = 4096;  // simulate W independent warps pulling 32-item chunks round-robin (order approximates the queue)
        RayS (*lanes)[32] = calloc(W, sizeof *lanes);
        char* open = malloc(W); memset(open, 1, W);
        int live = W;
        while (live) {
            for (int w = 0; w < W; w++) {
                if (!open[w]) continue;
                RayS* r = lanes[w];
                // one outer iteration of the kernel loop for warp w
                int ne = 0; for (int l = 0; l < 32; l++) ne += r[l].state == 0;
                if (ne >= R && next < n) {
                    cost += CREFILL + CGEN;
                    for (int l = 0; l < 32 && next < n; l++) if (r[l].state == 0) { ray_load(&r[l], seg + (size_t)next * max_seg, nseg[next]); next++; }
                }
                int any = 0; for (int l = 0; l < 32; l++) any |= r[l].state;
                if (!any) { if (next >= n) { open[w] = 0; live--; } continue; }
                for (;;) {
                    int ni = 0, nl = 0; for (int l = 0; l < 32; l++) { ni += r[l].state == 1; nl += r[l].state == 2; }
                    if (!ni) break;
                    if (ni < E && nl) break;
                    cost += CI + 4;
                    for (int l = 0; l < 32; l++) if (r[l].state == 1) ray_after_inner(&r[l]);
                }
                int maxt = 0; for (int l = 0; l < 32; l++) if (r[l].state == 2 && r[l].rem_t > maxt) maxt = r[l].rem_t;
                if (maxt) cost += CLEAF + (double)maxt * CT;
                for (int l = 0; l < 32; l++) if (r[l].state == 2) ray_after_leaf(&r[l]);
                cost += 6;
            }
        }
        printf("B  lanes refill>=%2d inner_exit<%2d        : warp-instr %.3e  efficiency %.3f\n", R, E, cost, useful / 32 / cost);
        free(lanes); free(open);
    }
    // ---- policy C: two rays per lane (slots a/b), inner phase picks a slot in inner state, leaf phase likewise
    for (int ei = 0; ei < 3; ei++) {
        int E = ei == 0 ? 24 : (ei == 1 ? 28 : 16);
        double cost = 0; long next = 0;
        const int W = 4096;
        RayS (*lanes)[64] = calloc(W, sizeof *lanes);
        char* open = malloc(W); memset(open, 1, W);
        int live = W;
        while (live) {
            for (int w = 0; w < W; w++) {
                if (!open[w]) continue;
                RayS* r = lanes[w];
                int ne = 0; for (int l = 0; l < 64; l++) ne += r[l].state == 0;
                if (ne >= 16 && next < n) {
                    cost += CREFILL + CGEN + CSWITCH;
                    int given = 0;
                    for (int l = 0; l < 64 && next < n && given < 32; l++) if (r[l].state == 0) { ray_load(&r[l], seg + (size_t)next * max_seg, nseg[next]); next++; given++; }
                }
                int any = 0; for (int l = 0; l < 64; l++) any |= r[l].state;
                if (!any) { if (next >= n) { open[w] = 0; live--; } continue; }
                // inner phase: lane l can work if slot l or l+32 is in inner state
                for (;;) {
                    int ni = 0, nl = 0;
                    for (int l = 0; l < 32; l++) { int a = r[l].state, b = r[l + 32].state; ni += (a == 1 || b == 1); nl += (a == 2 || b == 2) && !(a == 1 || b == 1); }
                    if (!ni) break;
                    if (ni < E && nl) break;
                    cost += CI + 6; g_isteps++; g_ilanes += ni;
                    int switches = 0;
                    for (int l = 0; l < 32; l++) {
                        if (r[l].state == 1) ray_after_inner(&r[l]);
                        else if (r[l + 32].state == 1) { ray_after_inner(&r[l + 32]); }
                    }
                    (void)switches;
                }
                // leaf phase: every lane with a slot in leaf state processes ONE leaf
                int maxt = 0, nl = 0;
                for (int l = 0; l < 32; l++) { RayS* p = r[l].state == 2 ? &r[l] : (r[l + 32].state == 2 ? &r[l + 32] : 0); if (p) { nl++; if (p->rem_t > maxt) maxt = p->rem_t; } }
                if (nl) { cost += CLEAF + (double)maxt * CT + CSWITCH; g_lphases++; g_llanes += nl; g_ltris += maxt; }
                for (int l = 0; l < 32; l++) { RayS* p = r[l].state == 2 ? &r[l] : (r[l + 32].state == 2 ? &r[l + 32] : 0); if (p) ray_after_leaf(p); }
                cost += 6;
            }
        }
        printf("C  two rays per lane, inner_exit<%2d      : warp-instr %.3e  efficiency %.3f\n", E, cost, useful / 32 / cost);
        printf("   inner steps %.3e avg lanes %.1f | leaf phases %.3e avg lanes %.1f avg max tris %.2f\n", g_isteps, g_ilanes / g_isteps, g_lphases, g_llanes / g_lphases, g_ltris / g_lphases);
        g_isteps = g_ilanes = g_lphases = g_llanes = g_ltris = 0;
        free(lanes); free(open);
    }
    // ---- policy D: K ray slots per lane, state in shared memory (inner step +10 instr, switching free);
    //      inner phase while >= E lanes have an inner-state slot; leaf phase tests ALL pending leaves of every lane
    int Ks[] = {2, 2, 2, 2, 3, 3, 3, 4, 4, 4}, EsD[] = {4, 8, 12, 16, 8, 12, 16, 8, 12, 16};
    for (int ci = 0; ci < 10; ci++) {
        int K = Ks[ci], E = EsD[ci];
        double cost = 0; long next = 0;
        const int W = 2048;
        RayS* all = calloc((size_t)W * 32 * K, sizeof(RayS));
        char* open = malloc(W); memset(open, 1, W);
        int live = W;
        while (live) {
            for (int w = 0; w < W; w++) {
                if (!open[w]) continue;
                RayS* r = all + (size_t)w * 32 * K;  // slot s of lane l = r[l*K+s]
                int ne = 0; for (int i = 0; i < 32 * K; i++) ne += r[i].state == 0;
                if (ne >= 32 && next < n) {  // refill one 32-item tile at a time
                    cost += CREFILL + CGEN + 20;
                    int given = 0;
                    for (int l = 0; l < 32 && next < n; l++)
                        for (int sl = 0; sl < K; sl++) if (r[l * K + sl].state == 0) { ray_load(&r[l * K + sl], seg + (size_t)next * max_seg, nseg[next]); next++; given++; break; }
                }
                int any = 0; for (int i = 0; i < 32 * K; i++) any |= r[i].state;
                if (!any) { if (next >= n) { open[w] = 0; live--; } continue; }
                for (;;) {
                    int ni = 0, nl = 0;
                    for (int l = 0; l < 32; l++) { int hi = 0, hl = 0; for (int sl = 0; sl < K; sl++) { hi |= r[l * K + sl].state == 1; hl |= r[l * K + sl].state == 2; } ni += hi; nl += (!hi && hl); }
                    if (!ni) break;
                    if (ni < E && nl) break;
                    cost += CI + 10 + 6; g_isteps++; g_ilanes += ni;
                    for (int l = 0; l < 32; l++) for (int sl = 0; sl < K; sl++) if (r[l * K + sl].state == 1) { ray_after_inner(&r[l * K + sl]); break; }
                }
                int maxt = 0, nl = 0, maxleaves = 0;
                for (int l = 0; l < 32; l++) { int t = 0, c = 0; for (int sl = 0; sl < K; sl++) if (r[l * K + sl].state == 2) { t += r[l * K + sl].rem_t; c++; } if (c) nl++; if (t > maxt) maxt = t; if (c > maxleaves) maxleaves = c; }
                if (nl) { cost += CLEAF + (double)maxt * CT + 20.0 * maxleaves; g_lphases++; g_llanes += nl; g_ltris += maxt; }
                for (int i = 0; i < 32 * K; i++) if (r[i].state == 2) ray_after_leaf(&r[i]);
                cost += 6;
            }
        }
        printf("D  %d slots/lane (smem state), inner_exit<%2d : warp-instr %.3e  efficiency %.3f\n", K, E, cost, useful / 32 / cost);
        printf("   inner steps %.3e avg lanes %.1f | leaf phases %.3e avg lanes %.1f avg max tris %.2f\n", g_isteps, g_ilanes / g_isteps, g_lphases, g_llanes / g_lphases, g_ltris / g_lphases);
        g_isteps = g_ilanes = g_lphases = g_llanes = g_ltris = 0;
        free(all); free(open);
    }
    return 0;
}

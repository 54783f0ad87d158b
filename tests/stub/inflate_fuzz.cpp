#include "inflate.h"
#include <zlib.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <string>
using namespace mkh;
static unsigned seed = 12345;
static unsigned rnd() { seed = seed * 1103515245u + 12345u; return (seed >> 8) & 0xffffff; }
static std::vector<uint8_t> make_raw() {
    std::vector<uint8_t> raw;
    int kind = rnd() % 5;
    size_t n = 1000 + rnd() % 300000;
    for (size_t i = 0; i < n; ++i) {
        if (kind == 0) raw.push_back("ACGT"[rnd() & 3]);
        else if (kind == 1) raw.push_back((uint8_t)rnd());
        else if (kind == 2) raw.push_back('A' + (i / 1000) % 3);
        else if (kind == 3) raw.push_back("ACGTN\n@+FFFF:,"[rnd() % 14]);
        else raw.push_back((uint8_t)(rnd() % 7 ? 'x' : rnd()));
    }
    return raw;
}
static std::vector<uint8_t> deflate_raw(const std::vector<uint8_t>& raw) {
    std::vector<uint8_t> comp(raw.size() + raw.size() / 8 + 1000);
    z_stream z{};
    int lv[] = {0, 1, 6, 9}; int st[] = {Z_DEFAULT_STRATEGY, Z_DEFAULT_STRATEGY, Z_FIXED, Z_HUFFMAN_ONLY, Z_RLE};
    deflateInit2(&z, lv[rnd() % 4], Z_DEFLATED, -15, 1 + rnd() % 9, st[rnd() % 5]);
    z.next_in = (Bytef*)raw.data(); z.avail_in = raw.size(); z.next_out = comp.data(); z.avail_out = comp.size();
    if (rnd() % 3 == 0) { z.avail_in = raw.size() / 2; deflate(&z, Z_SYNC_FLUSH); z.avail_in = raw.size() - raw.size() / 2; }
    deflate(&z, Z_FINISH);
    comp.resize(comp.size() - z.avail_out);
    deflateEnd(&z);
    return comp;
}
int main(int argc, char** argv) {
    int iters = argc > 1 ? atoi(argv[1]) : 300;
    seed = argc > 2 ? atoi(argv[2]) : 1;
    long n_ok = 0, n_err = 0;
    for (int it = 0; it < iters; ++it) {
        std::vector<uint8_t> raw = make_raw(), comp = deflate_raw(raw);
        for (int v = 0; v < 6; ++v) {
            std::vector<uint8_t> c = comp;
            if (v == 1) c.resize(rnd() % c.size());                       // cut
            if (v >= 2) for (int k = 0; k < 1 + (int)(rnd() % 4); ++k) c[rnd() % c.size()] ^= 1u << (rnd() % 8);  // flipped bits
            if (c.empty()) continue;
            // (1) byte decoder, exactly 16 readable bytes behind the input, output in a window like GzipStream's
            {
                std::vector<uint8_t> in(c.size() + 16);  // ASan: one byte more is an error
                memcpy(in.data(), c.data(), c.size());
                std::vector<uint8_t> out(raw.size() + 4096 + Inflater::kOutputMargin);
                Inflater inf;
                const uint8_t* ip = in.data(); uint8_t* op = out.data();
                Inflater::Status rc;
                size_t fed = std::min<size_t>(c.size(), 1 + rnd() % c.size());
                for (;;) {  // in pieces: NeedInput / OutputFull handling
                    bool fin = fed == c.size();
                    rc = inf.run(&ip, in.data() + fed, fin, out.data(), &op, out.data() + out.size() - 8);
                    if (rc == Inflater::kNeedInput && !fin) { fed = std::min(c.size(), fed + 1 + rnd() % 5000); continue; }
                    break;
                }
                if (v == 0 && (rc != Inflater::kStreamEnd || (size_t)(op - out.data()) != raw.size() || memcmp(out.data(), raw.data(), raw.size()))) { printf("VALID STREAM FAILED it %d rc %d\n", it, (int)rc); return 1; }
                (rc == Inflater::kStreamEnd ? n_ok : n_err)++;
            }
            // (2) exact one-shot (BGZF): output buffer of exactly the right size
            {
                std::vector<uint8_t> in(c.size() + 16); memcpy(in.data(), c.data(), c.size());
                std::vector<uint8_t> out(raw.size());
                bool ok = inflate_exact(in.data(), c.size(), out.data(), out.size());
                if (v == 0 && (!ok || memcmp(out.data(), raw.data(), raw.size()))) { printf("EXACT FAILED it %d\n", it); return 1; }
            }
            // (3) probe + marker decode from random bit positions, 600 readable bytes behind the input
            {
                std::vector<uint8_t> in(c.size() + 600); memcpy(in.data(), c.data(), c.size());
                Inflater inf; std::vector<uint16_t> sym;
                for (int k = 0; k < 20; ++k) {
                    uint64_t p = (uint64_t)(rnd() % c.size()) * 8 + rnd() % 8;
                    if (k == 0) p = 0;
                    if (k == 0 || inf.probe_dynamic_header(in.data(), in.data() + c.size(), p)) {
                        auto r = inf.run_markers(in.data(), in.data() + c.size(), p, p + 8 * (rnd() % 100000), &sym, 1 << 22);
                        if (r.ok && r.n_out > sym.size()) { printf("marker run overran its buffer\n"); return 1; }
                        if (v == 0 && k == 0 && r.ok) {  // from the first bit the window is never referenced: compare
                            size_t m = r.n_out - 32768; std::vector<uint8_t> win(32768), out(m);
                            if (!resolve_markers(sym.data() + 32768, m, win.data(), 0, out.data()) || memcmp(out.data(), raw.data(), m)) { printf("MARKER DECODE WRONG it %d\n", it); return 1; }
                        }
                    }
                }
                for (int k = 0; k < 2000; ++k) inf.probe_dynamic_header(in.data(), in.data() + c.size(), (uint64_t)(rnd() % c.size()) * 8 + rnd() % 8);
            }
        }
    }
    printf("done: %ld streams ended normally, %ld with an error\n", n_ok, n_err);
}

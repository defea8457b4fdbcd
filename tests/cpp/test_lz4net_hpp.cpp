// Exercises include/lz4net.hpp (the C++ host mirror of LZ4Codec / LZ4Stream) against the oracle port.
// Shape: src/LZ4.Tests/StreamTests.cs:47-63,148-181 (random-length writes, read back) and WrapTests.cs:11-48.
#include "lz4net.hpp"
#include "../../oracle/lz4_oracle.h"

#include <cstdio>
#include <cstdlib>
#include <sstream>

static uint32_t rng_state = 12345;
static uint32_t rnd() { rng_state = rng_state * 2654435761u + 2246822519u; return rng_state >> 8; }

static std::vector<uint8_t> make_data(size_t n, int kind)
{
    std::vector<uint8_t> d(n);
    for (size_t i = 0; i < n; i++) {
        if (kind == 0) d[i] = (uint8_t)rnd();                                         // incompressible
        else if (kind == 1) d[i] = (uint8_t)("lorem ipsum dolor sit amet "[(i * 7 + (rnd() % 3 == 0)) % 27]);   // text-like
        else d[i] = (i % 97 < 60) ? (uint8_t)(i % 251) : (uint8_t)rnd();              // mixed
    }
    return d;
}

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

int main()
{
    using namespace LZ4;
    Context ctx(0);
    // --- LZ4Codec: byte-identical to the oracle, round trips, error behaviour ---
    for (int kind = 0; kind < 3; kind++)
        for (size_t n : {size_t(1), size_t(13), size_t(2230), size_t(65536), size_t(200000)}) {
            auto d = make_data(n, kind);
            std::vector<uint8_t> want((size_t)lz4o_bound((int)n));
            auto enc = LZ4Codec::Encode(d);
            int w = lz4o_encode(d.data(), (int)n, want.data(), (int)want.size());
            CHECK((int)enc.size() == w && std::equal(enc.begin(), enc.end(), want.begin()));
            auto hc = LZ4Codec::EncodeHC(d);
            w = lz4o_encode_hc(d.data(), (int)n, want.data(), (int)want.size());
            CHECK((int)hc.size() == w && std::equal(hc.begin(), hc.end(), want.begin()));
            CHECK(LZ4Codec::Decode(enc, (int)n) == d && LZ4Codec::Decode(hc, (int)n) == d);
            std::vector<uint8_t> out(n);
            CHECK(LZ4Codec::Decode(enc.data(), 0, (int)enc.size(), out.data(), 0, (int)n, false) == (int)n && out == d);
            bool threw = false;
            try { LZ4Codec::Decode(enc.data(), 0, (int)enc.size() - 1, out.data(), 0, (int)n, true); } catch (const std::invalid_argument&) { threw = true; }
            CHECK(threw || n == 1);
            CHECK(LZ4Codec::Unwrap(LZ4Codec::Wrap(d, ctx), ctx) == d && LZ4Codec::Unwrap(LZ4Codec::WrapHC(d, ctx), ctx) == d);
        }
    {   // incompressible block with cap == n: fast -> 0, HC -> -1 (src/LZ4ps/LZ4Codec.Safe.cs:721-723)
        auto d = make_data(2048, 0); std::vector<uint8_t> out(2048);
        CHECK(LZ4Codec::Encode(d.data(), 0, 2048, out.data(), 0, 2048) == 0);
        CHECK(LZ4Codec::EncodeHC(d.data(), 0, 2048, out.data(), 0, 2048) == -1);
        CHECK(LZ4Codec::Encode(d.data(), 0, 0, out.data(), 0, 2048) == 0);
    }
    // --- LZ4Stream: random-length writes, then read back in random-length reads, both compression levels ---
    for (int hc = 0; hc < 2; hc++)
        for (int blockSize : {65536, 1000, 1 << 20}) {
            auto d = make_data(3 * 1000 * 1000 + 17, 2);
            std::stringstream inner;
            {
                LZ4Stream s((std::ostream&)inner, ctx, hc ? HighCompression : Default, blockSize, 8);
                size_t pos = 0;
                while (pos < d.size()) {
                    size_t k = std::min<size_t>(d.size() - pos, 1 + rnd() % 300000);
                    s.Write(d.data(), (int)pos, (int)k); pos += k;
                    if (rnd() % 5 == 0) s.Flush();
                }
                s.Close();
            }
            const std::string wire = inner.str();
            CHECK(lz4b200_stream_decoded_size(wire.data(), (int64_t)wire.size()) == (int64_t)d.size());
            for (int interactive = 0; interactive < 2; interactive++) {
                std::stringstream rd(wire);
                LZ4Stream r((std::istream&)rd, ctx, interactive ? InteractiveRead : Default, 5);
                std::vector<uint8_t> back; std::vector<uint8_t> buf(400000);
                for (;;) {
                    int want = 1 + (int)(rnd() % buf.size());
                    int got = r.Read(buf.data(), 0, want);
                    if (got == 0) break;
                    CHECK(interactive || got == want || back.size() + (size_t)got == d.size());
                    back.insert(back.end(), buf.begin(), buf.begin() + got);
                }
                CHECK(back == d);
            }
            // truncated stream -> EndOfStreamException analogue
            std::stringstream cut(wire.substr(0, wire.size() - 3));
            LZ4Stream r((std::istream&)cut, ctx);
            std::vector<uint8_t> buf(d.size() + 10); bool threw = false;
            try { while (r.Read(buf.data(), 0, (int)buf.size()) > 0) {} } catch (const std::exception&) { threw = true; }
            CHECK(threw);
        }
    std::printf("lz4net.hpp: all checks passed\n");
    return 0;
}

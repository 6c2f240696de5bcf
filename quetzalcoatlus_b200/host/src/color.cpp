// Host colour substrate: spectra evaluation, named spectra from the packed data file,
// the sRGB colour space, RGB -> spectrum table lookup, the pixel sensor, and the
// flatten() hooks that emit qz_spectrum records.  Scene-build time only.
//
// Reference behaviour followed (file:line into /root/reference/src/color):
//   spectrum.cpp:12-26 integral / inner_product; :47-53 dense lookup (lroundf);
//   :65-99 from_interleaved; :101-109 piecewise-linear lookup (quirks reproduced, see
//   spectrum.hpp); :113-133 blackbody; rgb.cpp:13-41 colour-space matrices; :51-67
//   sigmoid; :78-140 table lookup; :168-196 RGB spectra; sensor.cpp:22-89 sensor.
#include <dlfcn.h>

#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>

#include "../color/color.hpp"
#include "../util.hpp"

// ---------------------------------------------------------------- data directory
namespace qzhost {

static std::string g_data_dir;

static std::string default_data_dir() {
    if (const char* env = std::getenv("QZ_DATA_DIR")) return env;
    Dl_info info;
    if (dladdr(reinterpret_cast<const void*>(&default_data_dir), &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        auto slash = p.find_last_of('/');
        std::string dir = slash == std::string::npos ? "." : p.substr(0, slash);
        for (const char* rel : {"/data", "/../data", "/../../data"}) {
            std::string cand = dir + rel + "/spectra_tables.bin";
            if (FILE* f = std::fopen(cand.c_str(), "rb")) {
                std::fclose(f);
                return dir + rel;
            }
        }
    }
    return ".";
}

const std::string& data_dir() {
    if (g_data_dir.empty()) g_data_dir = default_data_dir();
    return g_data_dir;
}
void set_data_dir(const std::string& dir) { g_data_dir = dir; }

static const std::vector<float>& raw_table(const std::string& name) {
    static std::map<std::string, std::vector<float>> tables;
    static std::once_flag once;
    std::call_once(once, [] {
        std::string path = data_dir() + "/spectra_tables.bin";
        FILE* f = std::fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("quetzalcoatlus_b200: cannot open " + path + " (set QZ_DATA_DIR)");
        char magic[8];
        uint32_t n = 0;
        if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "QZSPEC01", 8) != 0 || std::fread(&n, 4, 1, f) != 1) {
            std::fclose(f);
            throw std::runtime_error("quetzalcoatlus_b200: bad spectra table file " + path);
        }
        for (uint32_t i = 0; i < n; i++) {
            char nm[25] = {0};
            uint32_t cnt = 0;
            if (std::fread(nm, 1, 24, f) != 24 || std::fread(&cnt, 4, 1, f) != 1) break;
            std::vector<float> v(cnt);
            if (std::fread(v.data(), 4, cnt, f) != cnt) break;
            tables[nm] = std::move(v);
        }
        std::fclose(f);
    });
    auto it = tables.find(name);
    if (it == tables.end()) throw std::runtime_error("quetzalcoatlus_b200: spectra table missing: " + name);
    return it->second;
}

}  // namespace qzhost

// ---------------------------------------------------------------- Spectrum base
float Spectrum::integral() const {
    float sum = 0.0f;
    for (int l = LAMBDA_MIN; l <= LAMBDA_MAX; ++l) sum += (*this)(l);
    return sum;
}

float Spectrum::inner_product(const Spectrum& other) const {
    float sum = 0.0f;
    for (int l = LAMBDA_MIN; l <= LAMBDA_MAX; ++l) sum += (*this)(l) * other(l);
    return sum;
}

int32_t ConstantSpectrum::flatten(qzhost::Flattener& f) const {
    int32_t id = f.find_spectrum(this);
    if (id >= 0) return id;
    qz_spectrum s{};
    s.kind = QZ_SPEC_CONSTANT;
    s.a = m_value;
    return f.add_spectrum(this, s);
}

// ---------------------------------------------------------------- densely sampled
DenselySampledSpectrum::DenselySampledSpectrum(std::vector<float>&& values, int lambda_min)
    : m_lambda_min(lambda_min), m_values(std::move(values)) {
    m_lambda_max = m_lambda_min + int(m_values.size()) - 1;
}

DenselySampledSpectrum::DenselySampledSpectrum(const Spectrum& other, int lambda_min, int lambda_max)
    : m_lambda_min(lambda_min), m_lambda_max(lambda_max) {
    m_values.resize(lambda_max - lambda_min + 1);
    for (size_t i = 0; i < m_values.size(); ++i) m_values[i] = other(lambda_min + int(i));
}

float DenselySampledSpectrum::operator()(float lambda) const {
    long index = std::lroundf(lambda - m_lambda_min);
    if (index < 0 || index >= long(m_values.size())) return 0.0f;
    return m_values[index];
}

int32_t DenselySampledSpectrum::flatten(qzhost::Flattener& f) const {
    int32_t id = f.find_spectrum(this);
    if (id >= 0) return id;
    qz_spectrum s{};
    s.kind = QZ_SPEC_DENSE;
    s.offset = f.add_pool(m_values.data(), m_values.size());
    s.count = uint32_t(m_values.size());
    s.aux = m_lambda_min;
    return f.add_spectrum(this, s);
}

// ---------------------------------------------------------------- piecewise linear
PiecewiseLinearSpectrum::PiecewiseLinearSpectrum(std::vector<float>&& lambdas, std::vector<float>&& values)
    : m_lambdas(std::move(lambdas)), m_values(std::move(values)) {
    if (m_lambdas.size() != m_values.size()) throw std::runtime_error("m_lambdas.size() != m_values.size()");
}

PiecewiseLinearSpectrum PiecewiseLinearSpectrum::from_interleaved(const std::vector<float>& in, bool normalize) {
    if (in.size() % 2 != 0) throw std::runtime_error("interleaved.size() % 2 != 0");
    if (in.size() == 0) throw std::runtime_error("interleaved.size() == 0");
    std::vector<float> lambdas, values;
    lambdas.reserve(in.size() / 2 + 2);
    values.reserve(in.size() / 2 + 2);
    if (in[0] > LAMBDA_MIN) {  // pad below
        lambdas.push_back(LAMBDA_MIN - 1);
        values.push_back(in[1]);
    }
    for (size_t i = 0; i < in.size(); i += 2) {
        lambdas.push_back(in[i]);
        values.push_back(in[i + 1]);
    }
    if (in.back() < LAMBDA_MAX) {  // tests the last VALUE, as the reference does
        lambdas.push_back(LAMBDA_MAX + 1);
        values.push_back(in.back());
    }
    PiecewiseLinearSpectrum spec(std::move(lambdas), std::move(values));
    if (normalize) {
        float c = spec.inner_product(*spectra::Y());
        for (float& v : spec.m_values) v = v * spectra::CIE_Y_INTEGRAL / c;
    }
    return spec;
}

float PiecewiseLinearSpectrum::operator()(float lambda) const {
    if (m_lambdas.empty() || lambda < m_lambdas.front() || lambda > m_lambdas.back()) return 0.0f;
    size_t n = m_lambdas.size();
    size_t i = size_t(std::lower_bound(m_lambdas.begin(), m_lambdas.end(), lambda) - m_lambdas.begin());
    // one-past-the-end reads are defined to be 0 (see header)
    float l1 = i + 1 < n ? m_lambdas[i + 1] : 0.0f;
    float v1 = i + 1 < n ? m_values[i + 1] : 0.0f;
    float t = (lambda - m_lambdas[i]) / (l1 - m_lambdas[i]);
    return m_values[i] * (1.0f - t) + v1 * t;
}

int32_t PiecewiseLinearSpectrum::flatten(qzhost::Flattener& f) const {
    int32_t id = f.find_spectrum(this);
    if (id >= 0) return id;
    qz_spectrum s{};
    s.kind = QZ_SPEC_PIECEWISE;
    s.count = uint32_t(m_lambdas.size());
    const float zero = 0.0f;
    s.offset = f.add_pool(m_lambdas.data(), m_lambdas.size());
    f.add_pool(&zero, 1);
    f.add_pool(m_values.data(), m_values.size());
    f.add_pool(&zero, 1);
    return f.add_spectrum(this, s);
}

// ---------------------------------------------------------------- blackbody
static float blackbody(float lambda, float t) {
    if (t <= 0.f) return 0.f;
    const float c = 299792458.f;
    const float h = 6.62606957e-34f;
    const float kb = 1.3806488e-23f;
    float l = lambda * 1e-9f;
    float le = 2.0f * h * c * c / (std::pow(lambda, 5) * (std::expm1f(h * c / (l * kb * t))));
    return le;
}

BlackbodySpectrum::BlackbodySpectrum(float t) : m_t(t) {
    float lambda_max = 2.897721e-12f / t;
    m_normalization_factor = 1.0f / blackbody(lambda_max, t);
}

float BlackbodySpectrum::operator()(float lambda) const { return m_normalization_factor * blackbody(lambda, m_t); }

int32_t BlackbodySpectrum::flatten(qzhost::Flattener& f) const {
    int32_t id = f.find_spectrum(this);
    if (id >= 0) return id;
    qz_spectrum s{};
    s.kind = QZ_SPEC_BLACKBODY;
    s.a = m_t;
    s.b = m_normalization_factor;
    return f.add_spectrum(this, s);
}

// ---------------------------------------------------------------- named spectra
namespace spectra {

static std::shared_ptr<const DenselySampledSpectrum> dense(const char* name) {
    return std::make_shared<DenselySampledSpectrum>(std::vector<float>(qzhost::raw_table(name)), 360);
}
static std::shared_ptr<const PiecewiseLinearSpectrum> piecewise(const char* name, bool normalize) {
    return std::make_shared<PiecewiseLinearSpectrum>(PiecewiseLinearSpectrum::from_interleaved(qzhost::raw_table(name), normalize));
}

#define QZ_NAMED_DENSE(fn, table) \
    std::shared_ptr<const DenselySampledSpectrum> fn() { static auto s = dense(table); return s; }
#define QZ_NAMED_PL(fn, table, norm) \
    std::shared_ptr<const PiecewiseLinearSpectrum> fn() { static auto s = piecewise(table, norm); return s; }

QZ_NAMED_DENSE(X, "CIE_X")
QZ_NAMED_DENSE(Y, "CIE_Y")
QZ_NAMED_DENSE(Z, "CIE_Z")
QZ_NAMED_PL(ILLUM_D65, "D65", true)
QZ_NAMED_PL(CANON_EOS_R, "CANON_R", false)
QZ_NAMED_PL(CANON_EOS_G, "CANON_G", false)
QZ_NAMED_PL(CANON_EOS_B, "CANON_B", false)
QZ_NAMED_PL(AL_IOR, "AL_IOR", false)
QZ_NAMED_PL(AL_ABSORPTION, "AL_ABSORPTION", false)
QZ_NAMED_PL(CU_IOR, "CU_IOR", false)
QZ_NAMED_PL(CU_ABSORPTION, "CU_ABSORPTION", false)
QZ_NAMED_PL(GLASS_BK7_IOR, "GLASS_BK7_IOR", false)
QZ_NAMED_PL(GLASS_SF11_IOR, "GLASS_SF11_IOR", false)

}  // namespace spectra

// ---------------------------------------------------------------- RGB <-> spectrum
static const size_t SPECTRUM_TABLE_RES = 32;

static float sigmoid(float x) {
    if (std::isinf(x)) return x > 0.0f ? 1.0f : 0.0f;
    return 0.5f + 0.5f * x / (std::sqrt(1.0f + x * x));
}

float RGBSigmoidPolynomial::operator()(float lambda) const { return sigmoid(c0 + c1 * lambda + c2 * lambda * lambda); }

float RGBSigmoidPolynomial::max_value() const {
    float result = std::max((*this)(LAMBDA_MIN), (*this)(LAMBDA_MAX));
    float lambda = -c1 / (2.0f * c0);
    if (lambda >= LAMBDA_MIN && lambda <= LAMBDA_MAX) result = std::max(result, (*this)(lambda));
    return result;
}

int32_t RGBSigmoidPolynomial::flatten(qzhost::Flattener& f) const {
    int32_t id = f.find_spectrum(this);
    if (id >= 0) return id;
    qz_spectrum s{};
    s.kind = QZ_SPEC_SIGMOID;
    s.a = c0; s.b = c1; s.c = c2;
    return f.add_spectrum(this, s);
}

std::shared_ptr<const RGBToSpectrumTable> RGBToSpectrumTable::sRGB() {
    static std::shared_ptr<const RGBToSpectrumTable> table;
    if (!table) {
        const size_t res = SPECTRUM_TABLE_RES;
        std::string path = qzhost::data_dir() + "/coeffs_SRGB_32.dat";
        FILE* f = std::fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("quetzalcoatlus_b200: cannot open " + path + " (set QZ_DATA_DIR)");
        std::vector<float> z(res), coeffs(3 * res * res * res * 3);
        bool ok = std::fread(z.data(), 4, z.size(), f) == z.size() &&
                  std::fread(coeffs.data(), 4, coeffs.size(), f) == coeffs.size();
        std::fclose(f);
        if (!ok) throw std::runtime_error("quetzalcoatlus_b200: short read on " + path);
        table = std::make_shared<RGBToSpectrumTable>(std::move(z), std::move(coeffs));
    }
    return table;
}

RGBSigmoidPolynomial RGBToSpectrumTable::operator()(const RGB& rgb) const {
    const size_t R = SPECTRUM_TABLE_RES;
    if (rgb.r() == rgb.g() && rgb.g() == rgb.b()) {  // grey: constant spectrum
        return RGBSigmoidPolynomial(0.0f, 0.0f, (rgb.r() - 0.5f) / std::sqrt(std::max(0.0f, rgb.r() * (1.0f - rgb.r()))));
    }
    const float comps[3] = {rgb.r(), rgb.g(), rgb.b()};
    size_t maxc = (comps[0] > comps[1]) ? ((comps[0] > comps[2]) ? 0 : 2) : ((comps[1] > comps[2]) ? 1 : 2);
    float z = comps[maxc];
    float x = comps[(maxc + 1) % 3] * (R - 1) / z;
    float y = comps[(maxc + 2) % 3] * (R - 1) / z;
    size_t xi = std::min(size_t(x), R - 2);
    size_t yi = std::min(size_t(y), R - 2);
    size_t zi = size_t(std::lower_bound(m_z_nodes.begin(), m_z_nodes.end(), z) - m_z_nodes.begin());
    if (zi != 0) zi--;
    if (zi > R - 2) {
        std::cout << "zi out of range, clamping to max" << std::endl;
        zi = R - 2;
    }
    float dx = x - xi, dy = y - yi;
    float dz = (z - m_z_nodes[zi]) / (m_z_nodes[zi + 1] - m_z_nodes[zi]);
    float c[3];
    for (size_t i = 0; i < 3; i++) {
        auto co = [&](size_t a, size_t b, size_t d) {
            return m_coeffs[maxc * R * R * R * 3 + (zi + d) * R * R * 3 + (yi + b) * R * 3 + (xi + a) * 3 + i];
        };
        c[i] = lerp(lerp(lerp(co(0, 0, 0), co(1, 0, 0), dx), lerp(co(0, 1, 0), co(1, 1, 0), dx), dy),
                    lerp(lerp(co(0, 0, 1), co(1, 0, 1), dx), lerp(co(0, 1, 1), co(1, 1, 1), dx), dy), dz);
    }
    return RGBSigmoidPolynomial(c[2], c[1], c[0]);
}

RGBColorSpace::RGBColorSpace(Vec2 r, Vec2 g, Vec2 b, std::shared_ptr<const Spectrum> illuminant,
                             std::shared_ptr<const RGBToSpectrumTable> table)
    : m_r(r), m_g(g), m_b(b), m_illuminant(illuminant), m_table(table) {
    XYZ white = XYZ::from_spectrum(*m_illuminant);
    m_white = white.xy();
    XYZ R = XYZ::from_xyY(r.x, r.y), G = XYZ::from_xyY(g.x, g.y), B = XYZ::from_xyY(b.x, b.y);
    Mat3 rgb({R.x, G.x, B.x, R.y, G.y, B.y, R.z, G.z, B.z});
    auto rgb_inv = rgb.invert();
    assert(rgb_inv.has_value());
    XYZ C = XYZ(rgb_inv.value() * white);
    m_xyz_from_rgb = rgb * Mat3::diagonal({C.x, C.y, C.z});
    auto inv = m_xyz_from_rgb.invert();
    assert(inv.has_value());
    m_rgb_from_xyz = inv.value();
}

RGBSigmoidPolynomial RGBColorSpace::to_spectrum(const RGB& rgb) const {
    return (*m_table)(RGB(std::clamp(rgb.x, 0.0f, 1.0f), std::clamp(rgb.y, 0.0f, 1.0f), std::clamp(rgb.z, 0.0f, 1.0f)));
}

std::shared_ptr<const RGBColorSpace> RGBColorSpace::sRGB() {
    static std::shared_ptr<const RGBColorSpace> space;
    if (!space) {
        space = std::make_shared<RGBColorSpace>(Vec2(0.64, 0.33), Vec2(0.3, 0.6), Vec2(0.15, 0.06),
                                                spectra::ILLUM_D65(), RGBToSpectrumTable::sRGB());
    }
    return space;
}

RGBUnboundedSpectrum::RGBUnboundedSpectrum(RGB rgb, const RGBColorSpace& cs) {
    m_scale = 2 * std::max({rgb.x, rgb.y, rgb.z});
    RGB scaled = m_scale ? RGB(rgb / m_scale) : RGB(0, 0, 0);
    m_polynomial = cs.to_spectrum(scaled);
}

int32_t RGBUnboundedSpectrum::flatten(qzhost::Flattener& f) const {
    int32_t id = f.find_spectrum(this);
    if (id >= 0) return id;
    qz_spectrum s{};
    s.kind = QZ_SPEC_RGB_UNBOUNDED;
    s.a = m_polynomial.c0; s.b = m_polynomial.c1; s.c = m_polynomial.c2;
    s.scale = m_scale;
    return f.add_spectrum(this, s);
}

RGBIlluminantSpectrum::RGBIlluminantSpectrum(RGB rgb, const RGBColorSpace& cs) : m_illuminant(cs.m_illuminant) {
    m_scale = 2 * std::max({rgb.x, rgb.y, rgb.z});
    RGB scaled = m_scale ? RGB(rgb / m_scale) : RGB(0, 0, 0);
    m_polynomial = cs.to_spectrum(scaled);
}

int32_t RGBIlluminantSpectrum::flatten(qzhost::Flattener& f) const {
    int32_t id = f.find_spectrum(this);
    if (id >= 0) return id;
    qz_spectrum s{};
    s.kind = QZ_SPEC_RGB_ILLUMINANT;
    s.a = m_polynomial.c0; s.b = m_polynomial.c1; s.c = m_polynomial.c2;
    s.scale = m_scale;
    s.aux = m_illuminant ? m_illuminant->flatten(f) : -1;
    return f.add_spectrum(this, s);
}

// ---------------------------------------------------------------- pixel sensor
static const float SENSOR_SATURATION = 40.0f;

static const Mat3 LMS_FROM_XYZ({0.8951, 0.2664, -0.1614, -0.7502, 1.7135, 0.0367, 0.0389, -0.0685, 1.0296});

// von-Kries style white balance; the result is stored but, as in the reference
// (sensor.cpp:19-31,38-40), never applied by to_sensor_rgb
static Mat3 white_balance(Vec2 source_white, Vec2 target_white) {
    auto src = LMS_FROM_XYZ * XYZ::from_xyY(source_white.x, source_white.y);
    auto dst = LMS_FROM_XYZ * XYZ::from_xyY(target_white.x, target_white.y);
    auto correct = Mat3::diagonal({dst.x / src.x, dst.y / src.y, dst.z / src.z});
    return LMS_FROM_XYZ * correct * LMS_FROM_XYZ;
}

PixelSensor::PixelSensor(const RGBColorSpace& cs, const Spectrum& illuminant, float imaging_ratio)
    : m_r(*spectra::X()), m_g(*spectra::Y()), m_b(*spectra::Z()), m_imaging_ratio(imaging_ratio) {
    m_xyz_from_sensor_rgb = white_balance(XYZ::from_spectrum(illuminant).xy(), cs.whitepoint());
}

PixelSensor::PixelSensor(const Spectrum& r, const Spectrum& g, const Spectrum& b, const RGBColorSpace& cs,
                         const Spectrum& illuminant, float imaging_ratio)
    : m_r(r), m_g(g), m_b(b), m_imaging_ratio(imaging_ratio) {
    m_xyz_from_sensor_rgb = white_balance(XYZ::from_spectrum(illuminant).xy(), cs.whitepoint());
}

RGB PixelSensor::to_sensor_rgb(const SpectrumSample& sample, const WavelengthSample& wl) const {
    auto l = sample / SpectrumSample::from_wavelengths_pdf(wl);
    RGB rgb((SpectrumSample::from_spectrum(m_r, wl) * l).average() * m_imaging_ratio,
            (SpectrumSample::from_spectrum(m_g, wl) * l).average() * m_imaging_ratio,
            (SpectrumSample::from_spectrum(m_b, wl) * l).average() * m_imaging_ratio);
    float m = std::max({rgb.x, rgb.y, rgb.z});
    if (m > SENSOR_SATURATION) rgb *= SENSOR_SATURATION / m;
    return rgb;
}

PixelSensor PixelSensor::CIE_XYZ(float imaging_ratio) {
    return PixelSensor(*RGBColorSpace::sRGB(), *spectra::ILLUM_D65(), imaging_ratio);
}

PixelSensor PixelSensor::CANON_EOS(float imaging_ratio) {
    return PixelSensor(*spectra::CANON_EOS_R(), *spectra::CANON_EOS_G(), *spectra::CANON_EOS_B(),
                       *RGBColorSpace::sRGB(), *spectra::ILLUM_D65(), imaging_ratio);
}

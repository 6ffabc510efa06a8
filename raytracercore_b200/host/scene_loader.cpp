// scene_loader.cpp — the reference's line-oriented scene text format (SceneLoader.cs:112-441, MatrixStack.cs,
// Raytracing/Objects/Cube.cs) restated for the host side of the B200 backend, so scenes authored for the
// reference load unchanged. Compile with -ffp-contract=off.
#include <cctype>
#include <cerrno>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <limits>
#include <sstream>

#include "scene.h"

namespace rtcore {
namespace {

struct MatrixStack {  // MatrixStack.cs
  std::vector<Mat4x4D> s{Mat4x4D::Identity()};
  const Mat4x4D& Peek() const {
    if (s.empty()) throw std::runtime_error("Stack empty.");
    return s.back();
  }
  void Push() { s.push_back(Peek()); }
  void Pop() {
    if (s.empty()) throw std::runtime_error("Stack empty.");
    s.pop_back();
  }
  void Transform(const Mat4x4D& m) {  // Push(Pop() * matrix)
    Mat4x4D t = Peek();
    Pop();
    s.push_back(t * m);
  }
  void InvTransform(const Mat4x4D& m) {  // Push(matrix * Pop())
    Mat4x4D t = Peek();
    Pop();
    s.push_back(m * t);
  }
};

// One line against lineRegex (SceneLoader.cs:38):
//   ^\s*(?:(\w+)(?:\s+([^\s,#]+)(?:\s*,?\s+([^\s,#]+))*)?)?\s*(?:#.*)?$
// Returns false when the line does not match. cmd is empty for blank / comment-only lines.
bool MatchLine(const std::string& line, std::string& cmd, std::vector<std::string>& params) {
  cmd.clear();
  params.clear();
  size_t n = line.size(), i = 0;
  auto is_ws = [](char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\f' || c == '\v'; };
  auto is_word = [](char c) { return std::isalnum((unsigned char)c) || c == '_'; };
  auto is_tok = [&](char c) { return !is_ws(c) && c != ',' && c != '#'; };
  auto tail_ok = [&](size_t p) {
    while (p < n && is_ws(line[p])) p++;
    return p == n || line[p] == '#';
  };
  while (i < n && is_ws(line[i])) i++;
  if (i == n || line[i] == '#') return true;
  size_t b = i;
  while (i < n && is_word(line[i])) i++;
  if (i == b) return false;
  cmd = line.substr(b, i - b);
  bool first = true;
  for (;;) {
    size_t save = i, ws1 = 0, ws2 = 0;
    bool comma = false;
    while (i < n && is_ws(line[i])) { i++; ws1++; }
    if (!first && i < n && line[i] == ',') {
      comma = true;
      i++;
      while (i < n && is_ws(line[i])) { i++; ws2++; }
    }
    bool sep_ok = comma ? ws2 >= 1 : ws1 >= 1;
    if (sep_ok && i < n && is_tok(line[i])) {
      size_t tb = i;
      while (i < n && is_tok(line[i])) i++;
      params.push_back(line.substr(tb, i - tb));
      first = false;
      continue;
    }
    return tail_ok(save);
  }
}

struct Params {
  const std::vector<std::string>* v;
  size_t pos = 0;
  std::string current;
  bool MoveNext() {
    if (pos >= v->size()) return false;
    current = (*v)[pos++];
    return true;
  }
  const std::string& Next() {  // SceneLoader.cs:42-48
    if (!MoveNext()) throw std::out_of_range("A parameter was missing from a command.");
    return current;
  }
};

// double.Parse(str, InvariantCulture), SceneLoader.cs:50-53. NumberStyles.Float | AllowThousands: [sign] digits with optional
// group separators, an optional fraction, an optional exponent -- or one of the symbols Infinity, -Infinity, NaN (matched
// without regard to case, as .NET Core 3.0+ does). strtod alone would also take "inf", "nan(...)", hex floats and a bare
// "infinity" prefix, which the reference rejects with a FormatException -> LoaderException.
double ParseDbl(const std::string& s) {
  const auto bad = []() { return std::invalid_argument("Input string was not in a correct format."); };
  if (s.empty()) throw bad();
  size_t i = 0;
  bool neg = false;
  if (s[i] == '+' || s[i] == '-') neg = s[i++] == '-';
  auto ieq = [&](const char* w) {
    size_t k = 0;
    for (; w[k]; k++)
      if (i + k >= s.size() || std::tolower((unsigned char)s[i + k]) != w[k]) return false;
    return i + k == s.size();
  };
  if (ieq("infinity")) return neg ? -std::numeric_limits<double>::infinity() : std::numeric_limits<double>::infinity();
  if (ieq("nan")) return std::numeric_limits<double>::quiet_NaN();
  std::string digits;  // the same number without group separators, for strtod
  if (neg) digits.push_back('-');
  size_t n_int = 0, n_frac = 0;
  while (i < s.size() && (std::isdigit((unsigned char)s[i]) || s[i] == ',')) {
    if (s[i] != ',') {
      digits.push_back(s[i]);
      n_int++;
    }
    i++;
  }
  if (i < s.size() && s[i] == '.') {
    digits.push_back(s[i++]);
    while (i < s.size() && std::isdigit((unsigned char)s[i])) {
      digits.push_back(s[i++]);
      n_frac++;
    }
  }
  if (n_int + n_frac == 0) throw bad();
  if (i < s.size() && (s[i] == 'e' || s[i] == 'E')) {
    digits.push_back(s[i++]);
    if (i < s.size() && (s[i] == '+' || s[i] == '-')) digits.push_back(s[i++]);
    size_t n_exp = 0;
    while (i < s.size() && std::isdigit((unsigned char)s[i])) {
      digits.push_back(s[i++]);
      n_exp++;
    }
    if (n_exp == 0) throw bad();
  }
  if (i != s.size()) throw bad();
  return std::strtod(digits.c_str(), nullptr);
}

int ParseInt(const std::string& s) {  // int.Parse, SceneLoader.cs:60-63 (NumberStyles.Integer: [sign] digits)
  if (s.empty()) throw std::invalid_argument("Input string was not in a correct format.");
  size_t i = (s[0] == '+' || s[0] == '-') ? 1 : 0;
  if (i == s.size()) throw std::invalid_argument("Input string was not in a correct format.");
  for (size_t k = i; k < s.size(); k++)
    if (!std::isdigit((unsigned char)s[k])) throw std::invalid_argument("Input string was not in a correct format.");
  errno = 0;
  long long v = std::strtoll(s.c_str(), nullptr, 10);
  if (errno == ERANGE || v > INT_MAX || v < INT_MIN) throw std::overflow_error("Value was either too large or too small for an Int32.");
  return (int)v;
}

// Cube (Raytracing/Objects/Cube.cs)
struct Cube {
  enum Side { XPos = 1, XNeg = 2, YPos = 4, YNeg = 8, ZPos = 16, ZNeg = 32, AllSides = 63 };
  Vec4D Position, Size;
  bool valid = false;

  static int GetSide(const std::string& name) {  // Cube.cs:22-63
    if (name == "implicit") return 0;
    if (name == "all") return AllSides;
    if (name.empty()) throw std::out_of_range("Index was outside the bounds of the array.");
    if (name[0] == '-' && name.size() == 2) {
      switch (name[1]) {
        case 'x': return XNeg;
        case 'y': return YNeg;
        case 'z': return ZNeg;
      }
    }
    char axis = ' ';
    if (name[0] == '+' && name.size() == 2)
      axis = name[1];
    else if (name.size() == 1)
      axis = name[0];
    switch (axis) {
      case 'x': return XPos;
      case 'y': return YPos;
      case 'z': return ZPos;
    }
    throw std::invalid_argument("Unknown Cube side name " + name + ".");
  }

  Primitive CreateRect(const Vec4D& pos, const Vec4D& up, const Vec4D& norm, double dist, double width, double height) const {
    // Cube.cs:71-76: Triangle.CreateRectangle(Ray.Directional(pos + norm * (dist / 2), up), norm, width, height)
    return Primitive::MakeRectangle(pos + (norm * (dist / 2)), Normalize(up), norm, width, height);
  }

  std::vector<Primitive> GetChildren(int sides) const {  // Cube.cs:90-116
    std::vector<Primitive> prims;
    if (sides & XPos) prims.push_back(CreateRect(Position, Vec4D(0, 0, 1, 0), Vec4D(1, 0, 0, 0), Size.X, Size.Y, Size.Z));
    if (sides & XNeg) prims.push_back(CreateRect(Position, Vec4D(0, 0, -1, 0), Vec4D(-1, 0, 0, 0), Size.X, Size.Y, Size.Z));
    if (sides & YPos) prims.push_back(CreateRect(Position, Vec4D(0, 0, 1, 0), Vec4D(0, 1, 0, 0), Size.Y, Size.X, Size.Z));
    if (sides & YNeg) prims.push_back(CreateRect(Position, Vec4D(0, 0, -1, 0), Vec4D(0, -1, 0, 0), Size.Y, Size.X, Size.Z));
    if (sides & ZPos) prims.push_back(CreateRect(Position, Vec4D(0, 1, 0, 0), Vec4D(0, 0, 1, 0), Size.Z, Size.X, Size.Y));
    if (sides & ZNeg) prims.push_back(CreateRect(Position, Vec4D(0, -1, 0, 0), Vec4D(0, 0, -1, 0), Size.Z, Size.X, Size.Y));
    return prims;
  }
};

struct VertexN {
  Vec4D pos, normal;
};

std::unique_ptr<Scene> Parse(std::istream& reader) {  // SceneLoader.FromFile, SceneLoader.cs:112-428
  std::unique_ptr<Scene> outScene(new Scene());

  // Camera state (:122-126)
  bool haveCam = false;
  Camera addCam;
  double imagePlane = 0, dofAmount = 0, focalLength = 0;
  Vec4D focalPoint;

  // Primitive state (:128-140)
  Cube obj;
  std::vector<Primitive> prims;
  bool twoSided = true, invert = false;
  const DoubleColor PH = DoubleColor::Placeholder();
  DoubleColor emission = PH, diffuse = PH, specular = PH, refraction = PH;
  double shininess = -1, refractionIndex = -1;

  MatrixStack stack, invStack;
  std::vector<Vec4D> vertices;
  std::vector<VertexN> verticesNormals;

  int lineNum = 1;
  std::string line, cmd;
  std::vector<std::string> pv;
  while (std::getline(reader, line)) {
    if (!MatchLine(line, cmd, pv)) throw std::runtime_error("Line did not match expected format.");  // :154-155
    if (!cmd.empty()) {
      for (char& c : cmd) c = (char)std::tolower((unsigned char)c);
      Params pe;
      pe.v = &pv;
      auto Next = [&]() -> const std::string& { return pe.Next(); };
      auto NextDbl = [&]() { return ParseDbl(pe.Next()); };
      auto NextInt = [&]() { return ParseInt(pe.Next()); };
      auto NextVec = [&](double w) {
        double x = NextDbl(), y = NextDbl(), z = NextDbl();
        return Vec4D(x, y, z, w);
      };
      auto Transf = [&](const Vec4D& v) { return stack.Peek() * v; };
      auto NextRGB = [&]() {
        double r = NextDbl(), g = NextDbl(), b = NextDbl();
        return DoubleColor(r, g, b);
      };
      auto NextBool = [&]() {  // :90-102
        const std::string& s = pe.Next();
        return s == "1" || s == "true" || s == "yes" || s == "y";
      };
      auto ReadAll = [&]() {
        std::vector<std::string> all;
        while (pe.MoveNext()) all.push_back(pe.current);
        return all;
      };
      auto Vertex = [&](std::vector<Vec4D>& list, int idx) -> Vec4D& {
        if (idx < 0 || (size_t)idx >= list.size()) throw std::out_of_range("Index was out of range.");
        return list[idx];
      };
      try {
        if (cmd == "size") {
          outScene->Width = NextInt();
          outScene->Height = NextInt();
        } else if (cmd == "background") {
          outScene->BackgroundRGB = NextRGB();
          outScene->BackgroundAlpha = NextDbl();
        } else if (cmd == "ambient") {
          const std::string& t = Next();
          if (t == "miss")
            outScene->AmbientRGB = PH;
          else if (t == "color")
            outScene->AmbientRGB = NextRGB();
          else
            throw std::runtime_error("Unknown ambient type " + t + ".");
        } else if (cmd == "recursion" || cmd == "bounce") {
          outScene->Recursion = NextInt();
        } else if (cmd == "debug") {
          const std::string& t = Next();
          if (t == "geom")
            outScene->DebugGeom = true;
          else if (t == "off")
            outScene->DebugGeom = false;
          else
            throw std::runtime_error("Unknown debug type " + t + ".");
        } else if (cmd == "dof") {  // :203-226
          imagePlane = NextDbl();
          dofAmount = NextDbl();
          const std::string& t = Next();
          if (t == "at") {
            focalPoint = Transf(NextVec(1));
            focalLength = 0;
          } else if (t == "to") {
            focalLength = NextDbl();
            focalPoint = Vec4D();
          } else if (t == "camera") {
            focalLength = 0;
            focalPoint = Vec4D();
          } else {
            throw std::runtime_error("Unknown dof focal command " + t + ".");
          }
        } else if (cmd == "camera" || cmd == "frustum" || cmd == "orthographic") {  // :227-240
          Vec4D pos = NextVec(1);
          Vec4D lookAt = NextVec(1);
          Vec4D up = Transf(NextVec(0) + pos);
          pos = Transf(pos);
          up = up - pos;
          addCam = Camera();
          addCam.initPosition = addCam.position = pos;
          addCam.initLookAt = addCam.lookAt = lookAt;
          addCam.initUp = addCam.up = up;
          if (cmd == "orthographic") {
            addCam.Kind = RTC_CAMERA_ORTHO;
            addCam.sizeMult = NextDbl();
          } else {
            addCam.Kind = RTC_CAMERA_FRUSTUM;
            addCam.fovY = toRadians(NextDbl());
          }
          haveCam = true;
        } else if (cmd == "twosided") {
          twoSided = NextBool();
        } else if (cmd == "invert") {
          invert = NextBool();
        } else if (cmd == "emission") {
          emission = NextRGB();
        } else if (cmd == "diffuse") {
          diffuse = NextRGB();
        } else if (cmd == "specular") {
          specular = NextRGB();
        } else if (cmd == "shininess") {  // :257-261
          shininess = NextDbl();
          if (pe.MoveNext()) shininess = std::pow(shininess, ParseDbl(pe.current));
        } else if (cmd == "refraction") {  // :262-273
          if (Next() == "off") {
            refraction = PH;
            refractionIndex = -1;
          } else {
            double r = ParseDbl(pe.current);
            double g = NextDbl(), b = NextDbl();
            refraction = DoubleColor(r, g, b);
            refractionIndex = NextDbl();
          }
        } else if (cmd == "translate") {  // :275-279
          Vec4D t = NextVec(0);
          stack.Transform(MatrixTransforms::Translate(t.X, t.Y, t.Z));
          invStack.InvTransform(MatrixTransforms::Translate(-t.X, -t.Y, -t.Z));
        } else if (cmd == "scale") {  // :280-284
          Vec4D s = NextVec(0);
          stack.Transform(MatrixTransforms::Scale(s.X, s.Y, s.Z));
          invStack.InvTransform(MatrixTransforms::Scale(1 / s.X, 1 / s.Y, 1 / s.Z));
        } else if (cmd == "rotate") {  // :285-290
          Vec4D axis = NextVec(0);
          double angle = NextDbl();
          stack.Transform(MatrixTransforms::Rotate(toRadians(angle), Normalize(axis)));
          invStack.InvTransform(MatrixTransforms::Rotate(-toRadians(angle), Normalize(axis)));
        } else if (cmd == "pushtransform") {
          stack.Push();
          invStack.Push();
        } else if (cmd == "poptransform") {
          stack.Pop();
          invStack.Pop();
        } else if (cmd == "sphere") {  // :300-302
          Vec4D c = NextVec(1);
          prims.push_back(Primitive::MakeSphere(c, NextDbl()));
        } else if (cmd == "plane") {  // :303-305
          double d = NextDbl();
          prims.push_back(Primitive::MakePlane(d, NextVec(0)));
        } else if (cmd == "vertex") {
          vertices.push_back(NextVec(1));
        } else if (cmd == "tri") {  // :309-320
          Vec4D p0 = Vertex(vertices, NextInt());
          Vec4D p1 = Vertex(vertices, NextInt());
          Vec4D p2 = Vertex(vertices, NextInt());
          bool mirror = false;
          if (pe.MoveNext() && pe.current == "mirrored") mirror = true;
          prims.push_back(Primitive::MakeTriangle(p0, p1, p2, mirror));
        } else if (cmd == "vertexnormal") {  // :321-323
          Vec4D p = NextVec(1);
          Vec4D nn = NextVec(0);
          verticesNormals.push_back(VertexN{p, nn});
        } else if (cmd == "trinormal") {  // :324-330
          auto get = [&](int idx) -> const VertexN& {
            if (idx < 0 || (size_t)idx >= verticesNormals.size()) throw std::out_of_range("Index was out of range.");
            return verticesNormals[idx];
          };
          const VertexN& a = get(NextInt());
          const VertexN& b = get(NextInt());
          const VertexN& c = get(NextInt());
          prims.push_back(Primitive::MakeTriangle(a.pos, a.normal, b.pos, b.normal, c.pos, c.normal));
        } else if (cmd == "cube") {  // :332-356
          Vec4D pos = NextVec(1);
          Vec4D size = NextVec(0);
          Cube cube;
          cube.Position = pos;
          cube.Size = size;
          cube.valid = true;
          obj = cube;
          if (pe.MoveNext()) {
            std::string opt = pe.current;
            int sides;
            if (opt == "all") {
              sides = Cube::AllSides;
            } else if (opt == "only") {
              sides = 0;
              for (const std::string& nme : ReadAll()) sides |= Cube::GetSide(nme);
            } else if (opt == "not") {
              sides = Cube::AllSides;
              for (const std::string& nme : ReadAll()) sides &= ~Cube::GetSide(nme);
            } else {
              throw std::runtime_error("Unknown option provided for cube construction: " + opt);
            }
            for (const Primitive& p : cube.GetChildren(sides)) prims.push_back(p);
          }
          // :355 obj.GetChildren(ImplicitInstance) -> Side 0 -> no primitives
        } else if (cmd == "instance") {  // :358-361
          for (const std::string& nme : ReadAll()) {
            if (!obj.valid) throw std::runtime_error("Object reference not set to an instance of an object.");
            for (const Primitive& p : obj.GetChildren(Cube::GetSide(nme))) prims.push_back(p);
          }
        } else if (cmd == "maxverts" || cmd == "maxvertnorms") {
        } else {
          // :367-369 unknown commands ("output", "point", "directional", ...) are logged and skipped
        }

        if (haveCam) {  // :372-386
          addCam.imagePlane = imagePlane;
          addCam.dofAmount = dofAmount;
          if (focalPoint != Vec4D())
            addCam.focalLength = Length(focalPoint - addCam.position);
          else if (focalLength != 0)
            addCam.focalLength = focalLength;
          else
            addCam.focalLength = Length(addCam.initLookAt - addCam.position);
          outScene->Cameras.push_back(addCam);
          haveCam = false;
        }

        for (Primitive& prim : prims) {  // :388-411
          prim.TwoSided = twoSided;
          prim.Invert = invert;
          if (emission != PH) prim.Emission = emission;
          if (diffuse != PH) prim.Diffuse = diffuse;
          if (specular != PH) prim.Specular = specular;
          if (shininess != -1) prim.Shininess = shininess;
          if (refraction != PH) {
            prim.Refraction = refraction;
            prim.RefractiveIndex = refractionIndex;
          }
          prim.Transform(stack.Peek(), invStack.Peek());
          outScene->AddPrimitive(prim);
        }
        prims.clear();
      } catch (const LoaderException&) {
        throw;
      } catch (const std::exception& e) {  // :415-421
        throw LoaderException(cmd, lineNum, e.what());
      }
    }
    lineNum++;
  }
  return outScene;
}

}  // namespace

namespace SceneLoader {
std::unique_ptr<Scene> FromFile(const std::string& filename) {
  std::ifstream f(filename);
  if (!f.good()) return nullptr;  // FileNotFoundException swallowed, SceneLoader.cs:430-439
  return Parse(f);
}
std::unique_ptr<Scene> FromString(const std::string& text) {
  std::istringstream f(text);
  return Parse(f);
}
}  // namespace SceneLoader

}  // namespace rtcore

// dataclass.h -- option classes with the user-visible behaviour of the reference's option bindings
// (R:helpers.h:40-283): attribute access, construction from a dict or from keyword arguments,
// recursive `mergedict`, `todict`, `summary`, copy / deepcopy, pickle, equality, and implicit
// dict -> Options conversion at call sites.
//
// Own design: every option class registers its fields explicitly (OptionsClass::field), and the generic
// members walk that registry through Python attribute access; the reference instead discovers
// attributes by introspecting dir().
#pragma once
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <sstream>
#include <string>
#include <vector>

namespace b2mh {
namespace py = pybind11;

template <typename T>
std::vector<std::string>& FieldNames() {
  static std::vector<std::string> names;
  return names;
}

inline std::string TypeNameOf(const py::handle& h) {
  return py::type::of(h).attr("__name__").cast<std::string>();
}

// Recursive update of `self` from `d`.  Unknown key -> AttributeError; value of the wrong type ->
// TypeError naming class, field, value and expected type (message format of R:helpers.h:87-121).
template <typename T>
void MergeDict(py::object self, const py::dict& d) {
  const std::string cls = TypeNameOf(self);
  for (const auto& item : d) {
    if (!py::isinstance<py::str>(item.first))
      throw py::type_error("Dictionary key is not a string: " + py::str(item.first).cast<std::string>());
    const std::string name = item.first.cast<std::string>();
    const auto& names = FieldNames<T>();
    if (std::find(names.begin(), names.end(), name) == names.end())
      throw py::attribute_error("'" + cls + "' object has no attribute '" + name + "'");
    py::object cur = self.attr(name.c_str());
    if (py::isinstance<py::dict>(item.second) && py::hasattr(cur, "mergedict")) {
      cur.attr("mergedict")(item.second);
      continue;
    }
    try {
      py::setattr(self, name.c_str(), item.second);
    } catch (py::error_already_set& e) {
      if (!e.matches(PyExc_TypeError)) throw;
      // the reference's fallback (R:helpers.h:70-84): coerce the value through the bases of the attribute's class
      // (an enum-like field given as int, a numpy scalar for a Python float, ...), then through the class itself
      bool coerced = false;
      py::object klass = cur.attr("__class__");
      py::list candidates = klass.attr("__bases__").cast<py::list>();
      candidates.append(klass);
      for (py::handle base : candidates) {
        try {
          py::setattr(self, name.c_str(), base(item.second));
          coerced = true;
          break;
        } catch (py::error_already_set&) {
        }
      }
      if (coerced) continue;
      std::ostringstream ss;
      ss << cls << "." << name << ": Could not convert " << TypeNameOf(item.second) << ": "
         << py::str(item.second).cast<std::string>() << " to '" << TypeNameOf(cur) << "'.";
      throw py::type_error("Failed to merge dict into class: Could not assign " + name + " (" + ss.str() + ")");
    }
  }
}

template <typename T>
py::dict ToDict(py::object self, bool recursive) {
  py::dict out;
  for (const std::string& name : FieldNames<T>()) {
    py::object v = self.attr(name.c_str());
    if (recursive && py::hasattr(v, "todict")) v = v.attr("todict")(recursive);
    out[py::str(name)] = v;
  }
  return out;
}

template <typename T>
std::string Summary(py::object self, bool write_type) {
  std::ostringstream ss;
  ss << TypeNameOf(self) << ":";
  for (const std::string& name : FieldNames<T>()) {
    py::object v = self.attr(name.c_str());
    ss << "\n    " << name;
    if (py::hasattr(v, "summary")) {
      std::istringstream sub(v.attr("summary")(write_type).cast<std::string>());
      std::string line;
      bool first = true;
      while (std::getline(sub, line)) {
        ss << (first ? ": " : "\n    ") << line;
        first = false;
      }
    } else {
      if (write_type) ss << ": " << TypeNameOf(v);
      ss << " = " << py::repr(v).cast<std::string>();
    }
  }
  return ss.str();
}

// py::class_ with a field registry; call Finish() after the last field().
template <typename T>
class OptionsClass : public py::class_<T> {
 public:
  OptionsClass(py::handle scope, const char* name, const char* doc = "") : py::class_<T>(scope, name, doc) {
    this->def(py::init<>());
  }
  template <typename M>
  OptionsClass& field(const char* name, M T::*member, const char* doc = "") {
    this->def_readwrite(name, member, doc);
    FieldNames<T>().push_back(name);
    return *this;
  }
  OptionsClass& Finish() {
    py::object cls = *this;
    {  // "<doc> (<type>, default: <value>)" on every field, like the reference's option classes (R:helpers.h:217-241)
      py::object defaults = cls();
      for (const std::string& name : FieldNames<T>()) {
        py::object member = defaults.attr(name.c_str());
        py::object prop = cls.attr(name.c_str());
        py::object old = prop.attr("__doc__");
        const std::string text = old.is_none() ? std::string() : py::str(old).cast<std::string>();
        const std::string value = py::hasattr(member, "summary") ? TypeNameOf(member) + "()" : py::str(member).cast<std::string>();
        prop.attr("__doc__") = py::str(text + (text.empty() ? "" : " ") + "(" + TypeNameOf(member) + ", default: " + value + ")");
      }
    }
    this->def(py::init([cls](const py::dict& d) {
                py::object self = cls();
                MergeDict<T>(self, d);
                return self.cast<T>();
              }),
              py::arg("dict"));
    this->def(py::init([cls](const py::kwargs& kw) {
      py::object self = cls();
      MergeDict<T>(self, py::dict(kw));
      return self.cast<T>();
    }));
    this->def("mergedict", [](py::object self, const py::dict& d) { MergeDict<T>(self, d); }, py::arg("dict"),
              "Recursively update the fields from a dict.");
    this->def("todict", [](py::object self, bool recursive) { return ToDict<T>(self, recursive); },
              py::arg("recursive") = true);
    this->def("summary", [](py::object self, bool write_type) { return Summary<T>(self, write_type); },
              py::arg("write_type") = false);
    this->def("__repr__", [](py::object self) { return Summary<T>(self, false); });
    this->def("__copy__", [](const T& self) { return T(self); });
    this->def("__deepcopy__", [](const T& self, const py::dict&) { return T(self); }, py::arg("memo"));
    this->def("__eq__", [](py::object self, py::object other) {
      return py::type::of(self).is(py::type::of(other)) && ToDict<T>(self, true).equal(ToDict<T>(other, true));
    });
    this->def(py::pickle([](py::object self) { return ToDict<T>(self, true); },
                         [cls](const py::dict& d) {
                           py::object self = cls();
                           MergeDict<T>(self, d);
                           return self.cast<T>();
                         }));
    py::implicitly_convertible<py::dict, T>();
    return *this;
  }
};

// Enums constructible from their member name (R:helpers.h:45-51); an unknown name raises IndexError
// (the reference throws std::out_of_range).
template <typename E>
void AddStringConstructor(py::enum_<E>& e) {
  py::object cls = e;
  e.def(py::init([cls](const std::string& name) {
    const py::dict members = cls.attr("__members__");
    if (!members.contains(py::str(name)))
      throw py::index_error("Invalid string value " + name + " for enum " + cls.attr("__name__").cast<std::string>());
    return members[py::str(name)].cast<E>();
  }));
  py::implicitly_convertible<std::string, E>();
}

}  // namespace b2mh

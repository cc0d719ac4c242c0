/*
 * MtlParser -- reference: source/MtlParser.{h,cpp}.  Same struct material_t, same class surface
 * (getMaterials, load), same keys incl. the custom ones (light, rough, p, nu, nv, Rs, Rd) and the
 * same quirks: lines shorter than 3 characters are skipped, `Tr` is ignored once ANY `d` has been
 * seen in the file (MtlParser.cpp:57, 99-108), values are read with atof.
 */
#ifndef MTLPARSER_H
#define MTLPARSER_H

#include <string>
#include <vector>

#include "cl_types.h"
#include "Logger.h"

using std::string;
using std::vector;

struct material_t {
	string mtlName;
	cl_float4 Ka;
	cl_float4 Kd;
	cl_float4 Ks;
	cl_float d;
	cl_float Ni;
	cl_float Ns;
	cl_char illum;
	// Light source yes/no
	cl_char light;
	// BRDF: Schlick
	cl_float rough;
	cl_float p;
	// BRDF: Shirley-Ashikhmin
	cl_float nu;
	cl_float nv;
	cl_float Rs;
	cl_float Rd;
};


class MtlParser {

	public:
		vector<material_t> getMaterials();
		void load( string file );
		/** Additive: install materials that were not read from a file (synthetic scenes). */
		void setMaterials( const vector<material_t>& materials );
		static material_t getEmptyMaterial();

	private:
		vector<material_t> mMaterials;

};

#endif

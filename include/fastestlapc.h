/* The reference's C API, exported by libfastestlap_b200.so -- same names, same argument lists as
 * /root/reference/src/main/c/fastestlapc.h:28-107, so that the reference's own front ends (src/main/python/fastest_lap.py binds
 * these symbols with ctypes, :10-14; the MATLAB binding likewise) load this library in place of libfastestlapc.
 *
 * Behind it: tables of named objects (vehicles, tracks, scalars, vectors: fastestlapc.cpp:30-58), the XML loaders of the vehicle
 * and track databases, and the GPU path -- the steady-state start point, the collocation NLP callbacks and the interior-point
 * solver of include/fastestlap_b200.h.  There is no CPU path: optimal_laptime and gg_diagram fail without a CUDA device.
 *
 * Error behaviour: the reference prints "[Fastest lap exception] -> <message>" and lets the C++ exception leave the C function
 * (fastestlapc.cpp:14-28).  Here the same line is printed and the function returns; fastestlapc_last_error() holds the message
 * (empty after a successful call).  Not offered (they fail with a message naming the reason): propagate_vehicle,
 * circuit_preprocessor, output channels other than the 45 of fl_nlp_output_name (vehicle_get_output and <output_variables> serve the tyre
 * slips / forces / dissipation / speeds and the chassis accelerations / forces / aerodynamics), and the optimal_laptime options compute_sensitivity, boost-time, and more than one hypermesh / constant control or integral constraint per call (one is
 * served: the F1's chassis.brake-bias as "constant" / "hypermesh", or one of engine-energy, tire-fl/-fr/-rl/-rr-energy). */
#ifndef FASTESTLAPC_B200_H
#define FASTESTLAPC_B200_H
#ifdef __cplusplus
extern "C" {
#endif

const char* fastestlapc_last_error(void);                                                              /* (addition) */

/* printing: fastestlapc.h:30-36 */
void set_print_level(int print_level);
void print_variables(void);
void print_variable(const char* variable_name);
void print_variable_to_string(char* str_out, const int n_char, const char* variable_name);
/* factories: fastestlapc.h:40-52 */
void create_vehicle_from_xml(const char* vehicle_name, const char* database_file);
void create_vehicle_empty(const char* vehicle_name, const char* vehicle_type);
void create_track_from_xml(const char* name, const char* track_file);
void create_vector(const char* name, const int n, double* data);
void create_scalar(const char* name, double value);
void copy_variable(const char* old_name, const char* new_name);
void move_variable(const char* old_name, const char* new_name);
/* destructors: fastestlapc.h:56 ("name*" deletes every variable with that prefix, "*" everything) */
void delete_variable(const char* variable_name);
/* getters: fastestlapc.h:60-80 */
void variable_type(char* variable_type, const int str_len_max, const char* variable_name);
double download_scalar(const char* name_c);
int download_vector_size(const char* name_c);
void download_vector(double* data, const int n, const char* name_c);
void vehicle_type_get_sizes(int* n_inputs, int* n_control, int* n_outputs, const char* c_vehicle_type_name);
void vehicle_type_get_names(char* key_name, char* input_names[], char* control_names[], char* output_names[], const int n_char, const char* vehicle_type_name);
double vehicle_get_output(const char* vehicle_name, const double* inputs, const double* controls, const double s, const char* property_name);
void vehicle_save_as_xml(const char* vehicle_name, const char* file_name);
int track_download_number_of_points(const char* track_name);
void track_download_data(double* data, const char* track_name, const int n, const char* variable_name_c);
double track_download_length(const char* track_name);
/* modifiers: fastestlapc.h:84-95 */
void vehicle_set_parameter(const char* vehicle_name, const char* parameter, const double value);
void vehicle_declare_new_constant_parameter(const char* c_vehicle_name, const char* parameter_path, const char* parameter_alias, const double parameter_value);
void vehicle_declare_new_variable_parameter(const char* c_vehicle_name, const char* parameter_path, const char* parameter_alias, const int n_parameters,
                                            const double* parameter_values, const int mesh_size, const int* mesh_parameter_indexes, const double* mesh_points);
void vehicle_change_track(const char* c_vehicle, const char* c_track);
void track_set_track_limit_correction(const char* c_track_name, const char* side, const int n_points, const double* s, const double* w_correction);
/* applications: fastestlapc.h:99-107 */
void steady_state(double* input_states, double* controls, const char* vehicle_name, double v, double ax, double ay);
void propagate_vehicle(double* input_states, double* controls, const char* vehicle_name, const char* track_name, double s, double ds, double* u_next, int use_circuit, const char* options);
void gg_diagram(double* ay, double* ax_max, double* ax_min, const char* vehicle_name, double v, const int n_points);
void optimal_laptime(const char* c_vehicle, const char* c_track_name, const int n_points, const double* s, const char* options);
void circuit_preprocessor(const char* options);

#ifdef __cplusplus
}
#endif
#endif
